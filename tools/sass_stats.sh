#!/bin/bash
# Static SASS statistics of stream_kernel<false,false,0> in a built library (proxy for the dynamic mix while off-GPU).
#   tools/sass_stats.sh path/to/lib.so [top_n]
lib=$1
cuobjdump -sass -fun '_ZN4bump13stream_kernelILb0ELb0ELi0EEEvNS_7ColumnsENS_4WorkEPKiPKdPdPy' "$lib" > /tmp/sass_$$.txt 2>/dev/null
tot=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass_$$.txt)
fp=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+\s+)?(DFMA|DMUL|DADD|DSETP|DMNMX)" /tmp/sass_$$.txt)
echo "total $tot fp64 $fp other $((tot-fp))"
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass_$$.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${2:-30} | awk '{printf "%s:%s ", $2, $1} END{print ""}'
rm -f /tmp/sass_$$.txt

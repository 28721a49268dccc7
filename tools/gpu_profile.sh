#!/bin/bash
# Run under gpurun: plain bench, then the ncu launch list and one full capture of the streaming kernel.
#   tools/gpu_profile.sh <tag> [bench args...]
set -u
tag=$1; shift
out=gpurun_out
cmd="python bench.py --steps 5 --warmup 3 --no-cpu-baseline $*"
$cmd > $out/plain_$tag.json 2> $out/plain_$tag.err || { echo "plain run failed"; tail -5 $out/plain_$tag.err; exit 1; }
cut -c1-400 $out/plain_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_list_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -f -o $out/prof_$tag $cmd > $out/ncu_full_$tag.log 2>&1
echo "full capture rc=$?"

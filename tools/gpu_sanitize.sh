#!/bin/bash
# compute-sanitizer over every kernel of the hot path (run under gpurun); summaries -> gpurun_out/r02_sanitizer_*.txt
set -u
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $out/r02_sanitizer_$tool.txt 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" $out/r02_sanitizer_$tool.txt | head -8
done

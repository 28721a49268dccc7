#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python tools/tune.py 2>&1 | tail -1 | tee -a $out/r12_tune.txt
bash tools/gpu_run9.sh r12

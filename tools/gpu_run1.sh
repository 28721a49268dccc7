#!/bin/bash
# GPU run 1 (round 2): kernel variants A/B, the gpu test suite, the bench line.
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r1_smi.txt 2>&1
echo "== tune" 
for v in default e10r2 e8r4 divw newton2; do
  if [ $v = default ]; then lib=bumpcosmology_b200/libbump_b200.so; else lib=build/libbump_$v.so; fi
  BUMP_LIB_PATH=$PWD/$lib timeout 300 python tools/tune.py 2>&1 | tail -1 | tee -a $out/r1_tune.txt
done
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $out/r1_smoke.txt
echo "== pytest"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r1_pytest.txt
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r1_bench.json 2> $out/r1_bench.err; echo "bench rc=$?"; cut -c1-1500 $out/r1_bench.json; tail -5 $out/r1_bench.err

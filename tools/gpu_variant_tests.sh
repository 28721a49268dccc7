#!/bin/bash
# The GPU tests that exercise concurrency / changing theta, once per library build, then the A/B timing:
#   tools/gpu_variant_tests.sh tag variant1 variant2 ...
set -u
out=gpurun_out; mkdir -p $out
tag=$1; shift
for v in "$@"; do
  if [ $v = default ]; then lib=bumpcosmology_b200/libbump_b200.so; else lib=build/libbump_$v.so; fi
  echo "== $v" | tee -a $out/${tag}_tests.txt
  BUMP_LIB_PATH=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_device_entry.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5 | tee -a $out/${tag}_tests.txt
done
bash tools/gpu_ab.sh $tag "$@"

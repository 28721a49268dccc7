#!/bin/bash
# parity tests against one variant build, then the A/B timing: tools/gpu_ab_parity.sh tag variant [others...]
set -u
out=gpurun_out; mkdir -p $out
tag=$1; shift
BUMP_LIB_PATH=$PWD/build/libbump_$1.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -4 | tee $out/${tag}_pytest.txt
bash tools/gpu_ab.sh $tag default "$@"

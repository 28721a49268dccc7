#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
tag=$1; shift
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $out/${tag}_pytest.txt
bash tools/gpu_ab.sh $tag "$@"

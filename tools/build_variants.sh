#!/bin/bash
# Build tuning variants of the library into build/ (git-ignored; travels to the GPU box with gpurun).
#   tools/build_variants.sh name1 "flags1" name2 "flags2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  echo "== build/libbump_$name.so  [$extra]"
  nvcc $FLAGS $extra -shared -o build/libbump_$name.so bumpcosmology_b200/csrc/bump_lib.cu bumpcosmology_b200/csrc/bump_nuts.cpp -ldl &
done
wait
ls -la build/*.so

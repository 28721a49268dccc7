// Compile-only harness: one instantiation of the streaming kernel, for quick SASS experiments
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -cubin -o build/dev.cubin tools/dev_stream.cu [-D...]
//   python tools/sass_loop.py build/dev.cubin
#include "../bumpcosmology_b200/csrc/bump_stream.cuh"
template __global__ void bump::stream_kernel<false, false, 0>(const bump::Columns, const bump::Work, const int*,
                                                              const double*, double*, unsigned long long*);

// Microbenchmark: issue cost of FP64-pipe instructions on B200 as a function of where their operands come from.
// 8 independent chains per thread, 16 warps per SM (4 per sub-partition), cycles per warp-instruction per
// sub-partition.  Variants:
//   0  x = fma(x, y_c, z_c)        three distinct vector registers
//   1  x = fma(x, y_c, K)          two registers + constant bank
//   2  x = fma(x, y, z_c)          y shared by all chains (operand-reuse cache candidate)
//   3  x = fma(x, y_c, x)          three register slots, two distinct registers
//   4  x = x * y_c                 DMUL, two registers
//   5  x = x + y_c                 DADD, two registers
//   6  x = fma(x, K1, K2)          one register (K1 constant bank, K2 uniform register / immediate)
//   7  x = fma(y_c, z_c, x) then x = fma(x, K, K2): alternating 3-register and 1-register
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu && ./fp64_operands
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 2048;
__constant__ double K[4] = {0.999999, 1e-9, 1.0000001, 3e-10};

template <int V, int NI = 0>
__global__ void __launch_bounds__(512, 1) op_kernel(double* out, double a, double b, long long* cycles) {
    double x[8], y[8], z[8];
    int n[NI > 0 ? NI : 1];
#pragma unroll
    for (int c = 0; c < NI; ++c) n[c] = threadIdx.x + c;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        x[c] = threadIdx.x * 1e-3 + c;
        y[c] = a + 1e-9 * (threadIdx.x & (c + 1));
        z[c] = b + 1e-12 * (threadIdx.x & (c + 3));
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (V == 0) x[c] = fma(x[c], y[c], z[c]);
            if (V == 1) x[c] = fma(x[c], y[c], K[1]);
            if (V == 2) x[c] = fma(x[c], y[0], z[c]);
            if (V == 3) x[c] = fma(x[c], y[c], x[c]);
            if (V == 4) x[c] = x[c] * y[c];
            if (V == 5) x[c] = x[c] + z[c];
            if (V == 6) x[c] = fma(x[c], K[0], b);
            if (c < NI) n[c] = (n[c] + 12345) ^ 0x5a5a;   // two ALU instructions with one register operand each
        }
        if (V == 7) {
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = fma(fma(y[c], z[c], x[c]), K[0], b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    int m = 0;
#pragma unroll
    for (int c = 0; c < NI; ++c) m += n[c];
    if (s == 12345.678 || m == 123456789) out[0] = s + m;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int V, int NI = 0>
void run(const char* what, int n_fp64, double* out, long long* d_cyc, int sms) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) op_kernel<V, NI><<<sms, 512>>>(out, 0.999999, 1e-9, d_cyc);
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double cyc = (double)c / ITERS / 4.0;   // 4 warps per sub-partition
    printf("{\"variant\": \"%s\", \"fp64_inst_per_iter\": %d, \"alu_inst_per_iter\": %d, "
           "\"cycles_per_iter_per_sp\": %.2f, \"cycles_per_fp64_inst_per_sp\": %.3f}\n",
           what, n_fp64, 2 * NI, cyc, cyc / n_fp64);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    double* out;
    long long* d_cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&d_cyc, 8);
    const int sms = p.multiProcessorCount;
    run<0>("DFMA R,R,R", 8, out, d_cyc, sms);
    run<1>("DFMA R,R,c[]", 8, out, d_cyc, sms);
    run<2>("DFMA R,Rshared,R", 8, out, d_cyc, sms);
    run<3>("DFMA Ra,Rb,Ra", 8, out, d_cyc, sms);
    run<4>("DMUL R,R", 8, out, d_cyc, sms);
    run<5>("DADD R,R", 8, out, d_cyc, sms);
    run<6>("DFMA R,c[],UR", 8, out, d_cyc, sms);
    run<7>("DFMA R,R,R + DFMA R,c[],UR", 16, out, d_cyc, sms);
    // co-issue: 8 FP64 instructions + 8 / 16 single-register ALU instructions (IADD imm, LOP3 imm) per iteration
    run<0, 4>("DFMA R,R,R + 8 ALU", 8, out, d_cyc, sms);
    run<0, 8>("DFMA R,R,R + 16 ALU", 8, out, d_cyc, sms);
    run<1, 4>("DFMA R,R,c[] + 8 ALU", 8, out, d_cyc, sms);
    run<1, 8>("DFMA R,R,c[] + 16 ALU", 8, out, d_cyc, sms);
    run<4, 4>("DMUL R,R + 8 ALU", 8, out, d_cyc, sms);
    run<4, 8>("DMUL R,R + 16 ALU", 8, out, d_cyc, sms);
    run<6, 4>("DFMA R,c[],UR + 8 ALU", 8, out, d_cyc, sms);
    run<6, 8>("DFMA R,c[],UR + 16 ALU", 8, out, d_cyc, sms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

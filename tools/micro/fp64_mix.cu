// Microbenchmark: what does the FP64 pipe sustain when DFMAs share the issue port / register file with other
// instructions?  The streaming kernel executes 256.5 FP64-pipe + 284 other warp instructions per warp-sample and
// sits at 58 % of the FP64 pipe whether 3 or 4 warps per sub-partition are resident, i.e. it is not latency bound.
// Each thread runs 8 independent DFMA chains interleaved with NI independent integer chains.
//   REGS = false: x = fma(x, a, b) with warp-uniform a, b  (DFMA R, R, UR, UR: one 64-bit register-file read)
//   REGS = true : x = fma(x, y_c, z_c) with per-thread y_c, z_c (three 64-bit register-file reads)
//   IOP = 0: n = n * ia + ib (IMAD, three register reads when REGS)   IOP = 1: n = (n + 12345) ^ 0x5a5a (two
//   single-operand ALU instructions: IADD/LOP3 with immediates)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu && ./fp64_mix
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 2048;

template <int NI, bool REGS, int IOP>
__global__ void __launch_bounds__(512, 1) mix_kernel(double* out, double a, double b, int ia, int ib, long long* cycles) {
    double x[8], y[8], z[8];
    int n[NI > 0 ? NI : 1];
    int ja = ia, jb = ib;
    if (REGS) {   // per-thread values: the compiler cannot keep them in uniform registers
        ja += threadIdx.x & 1;
        jb += threadIdx.x & 2;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        x[c] = threadIdx.x * 1e-3 + c;
        y[c] = a + (REGS ? 1e-9 * (threadIdx.x & (c + 1)) : 0.0);
        z[c] = b + (REGS ? 1e-12 * (threadIdx.x & (c + 3)) : 0.0);
    }
#pragma unroll
    for (int c = 0; c < NI; ++c) n[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            x[c] = fma(x[c], y[c], z[c]);
#pragma unroll
            for (int k = c; k < NI; k += 8) {
                if (IOP == 0) n[k] = n[k] * ja + jb;
                else n[k] = (n[k] + 12345) ^ 0x5a5a;
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
    int m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
#pragma unroll
    for (int c = 0; c < NI; ++c) m += n[c];
    if (s == 12345.678 || m == 123456789) out[0] = s + m;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NI, bool REGS, int IOP>
void run(int warps, double* out, long long* d_cyc, int sms) {
    long long c = 0;
    mix_kernel<NI, REGS, IOP><<<sms, warps * 32>>>(out, 0.999999, 1e-9, 3, 7, d_cyc);
    mix_kernel<NI, REGS, IOP><<<sms, warps * 32>>>(out, 0.999999, 1e-9, 3, 7, d_cyc);
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double other = (IOP == 0 ? NI : 2.0 * NI) + 3.0;   // + loop counter, compare, branch
    const double wps = warps / 4.0;                            // warps per sub-partition
    const double cyc = (double)c / ITERS / wps;                // cycles of a sub-partition per warp-iteration
    printf("{\"warps\": %d, \"dfma_operands\": \"%s\", \"int_op\": \"%s\", \"fp64_inst\": 8, \"other_inst\": %.0f, "
           "\"cycles_per_warp_iter_per_sp\": %.2f, \"fp64_pipe_frac\": %.3f, \"issue_frac\": %.3f}\n",
           warps, REGS ? "R,R,R" : "R,UR,UR", IOP == 0 ? (REGS ? "IMAD R,R,R" : "IMAD R,UR,UR") : "IADD imm + LOP3 imm",
           other, cyc, 16.0 / cyc, (8.0 + other) / cyc);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    double* out;
    long long* d_cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&d_cyc, 8);
    const int sms = p.multiProcessorCount;
    for (int warps : {12, 16}) {
        run<0, false, 0>(warps, out, d_cyc, sms);
        run<4, false, 0>(warps, out, d_cyc, sms);
        run<8, false, 0>(warps, out, d_cyc, sms);
        run<4, false, 1>(warps, out, d_cyc, sms);
        run<0, true, 0>(warps, out, d_cyc, sms);
        run<4, true, 0>(warps, out, d_cyc, sms);
        run<8, true, 0>(warps, out, d_cyc, sms);
        run<2, true, 1>(warps, out, d_cyc, sms);
        run<4, true, 1>(warps, out, d_cyc, sms);
        run<6, true, 1>(warps, out, d_cyc, sms);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

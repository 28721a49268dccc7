// Microbenchmark for the mixed-precision question: what would an fp32 version of the
// streaming kernel pay per instruction on B200?  Same harness as fp64_operands.cu: 16 warps per SM, 8 independent
// chains per thread, cycles per warp-instruction per SM sub-partition.
//   0  FFMA R,R,R            three distinct registers (the fp64 version of this issues every 3 cycles)
//   1  FFMA R,R,c[]
//   2  FFMA R,R,R + 8 ALU    co-issue with single-register integer instructions
//   3  MUFU.EX2              ex2.approx.ftz.f32
//   4  MUFU.RCP              rcp.approx.ftz.f32
//   5  MUFU.LG2              lg2.approx.ftz.f32
//   6  FFMA + MUFU.EX2       one exp per 8 FFMAs (the kernel's rough ratio)
//   7  F2F.F64.F32 + DADD    fp32 product accumulated in fp64 (one conversion + one add per value)
//   8  Kahan fp32 add        compensated accumulation (4 FADD) as the alternative to 7
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_mix fp32_mix.cu && ./fp32_mix
// Results: profiles/r01b_fp32_issue_microbench.txt.
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITERS = 2048;
__constant__ float KF[4] = {0.999999f, 1e-6f, 1.0000001f, 3e-7f};

template <int V>
__global__ void __launch_bounds__(512, 1) f32_kernel(float* out, float a, float b, long long* cycles) {
    float x[8], y[8], z[8], c[8];
    double acc[8];
    int n[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = threadIdx.x * 1e-3f + k + 1.0f;
        y[k] = a + 1e-6f * (threadIdx.x & (k + 1));
        z[k] = b + 1e-7f * (threadIdx.x & (k + 3));
        c[k] = 0.0f;
        acc[k] = 0.0;
        n[k] = threadIdx.x + k;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (V == 0 || V == 2 || V == 6) x[k] = fmaf(x[k], y[k], z[k]);
            if (V == 1) x[k] = fmaf(x[k], y[k], KF[1]);
            if (V == 2) n[k] = (n[k] + 12345) ^ 0x5a5a;
            if (V == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
            if (V == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
            if (V == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
            if (V == 6 && k == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y[0]));
            if (V == 7) {
                x[k] = x[k] * y[k];
                acc[k] += (double)x[k];
            }
            if (V == 8) {   // Kahan: c carries the rounding error of the running sum z
                x[k] = x[k] * y[k];
                const float t = x[k] - c[k];
                const float s = z[k] + t;
                c[k] = (s - z[k]) - t;
                z[k] = s;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
    int m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s += x[k] + y[k] + z[k] + c[k] + (float)acc[k];
        m += n[k];
    }
    if (s == 12345.678f || m == 123456789) out[0] = s + m;
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int V>
void run(const char* what, int n_inst, float* out, long long* d_cyc, int sms) {
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) f32_kernel<V><<<sms, 512>>>(out, 0.999999f, 1e-6f, d_cyc);
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double cyc = (double)c / ITERS / 4.0;   // 4 warps per sub-partition
    printf("{\"variant\": \"%s\", \"counted_inst_per_iter\": %d, \"cycles_per_iter_per_sp\": %.2f, "
           "\"cycles_per_counted_inst_per_sp\": %.3f}\n",
           what, n_inst, cyc, cyc / n_inst);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    float* out;
    long long* d_cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&d_cyc, 8);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA R,R,R", 8, out, d_cyc, sms);
    run<1>("FFMA R,R,c[]", 8, out, d_cyc, sms);
    run<2>("FFMA R,R,R + 16 ALU (count = FFMA)", 8, out, d_cyc, sms);
    run<3>("MUFU.EX2", 8, out, d_cyc, sms);
    run<4>("MUFU.RCP", 8, out, d_cyc, sms);
    run<5>("MUFU.LG2", 8, out, d_cyc, sms);
    run<6>("8 FFMA + 1 MUFU.EX2 (count = 9)", 9, out, d_cyc, sms);
    run<7>("FMUL + F2F.F64.F32 + DADD (count = values)", 8, out, d_cyc, sms);
    run<8>("FMUL + Kahan 4 FADD (count = values)", 8, out, d_cyc, sms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

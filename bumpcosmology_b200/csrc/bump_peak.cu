// FP64 issue-rate probe: the roofline denominator for the fp64 hyperlikelihood path (MEASURED_PEAKS.json has
// HBM and bf16 figures only).  Prints one JSON line: DFMA warp-instructions are counted per thread.
//   bump_peak [device]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int CHAINS = 8, ITERS = 4096, UNROLL = 16;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < ITERS / UNROLL; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;   // never true; keeps the chains alive
}

// dependent-issue latency: one chain per thread, one warp per SM sub-partition
template <int NCH>
__global__ void __launch_bounds__(128) dfma_latency_kernel(double* out, double a, double b, long long* cycles) {
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) x[c] = fma(x[c], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main(int argc, char** argv) {
    int dev = argc > 1 ? atoi(argv[1]) : 0;
    if (cudaSetDevice(dev) != cudaSuccess) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    double* out;
    cudaMalloc(&out, 8);
    const int blocks = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, 256>>>(out, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) {
        printf("{\"error\": \"kernel failed\"}\n");
        return 1;
    }
    long long* d_cyc;
    cudaMalloc(&d_cyc, 8);
    long long c1 = 0, c2 = 0, c4 = 0;
    dfma_latency_kernel<1><<<p.multiProcessorCount, 128>>>(out, 0.999999, 1e-9, d_cyc);
    cudaMemcpy(&c1, d_cyc, 8, cudaMemcpyDeviceToHost);
    dfma_latency_kernel<2><<<p.multiProcessorCount, 128>>>(out, 0.999999, 1e-9, d_cyc);
    cudaMemcpy(&c2, d_cyc, 8, cudaMemcpyDeviceToHost);
    dfma_latency_kernel<4><<<p.multiProcessorCount, 128>>>(out, 0.999999, 1e-9, d_cyc);
    cudaMemcpy(&c4, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)blocks * 256 * CHAINS * ITERS;   // thread-level DFMAs
    const double rate = n / (best * 1e-3);
    printf("{\"dfma_per_s\": %.6e, \"fp64_tflops\": %.3f, \"sms\": %d, \"ms\": %.4f, \"dfma_per_clk_per_sm_at_max_clock\": %.2f, "
           "\"max_clock_mhz\": %d, \"dfma_dependent_latency_cycles\": %.2f, \"cycles_per_dfma_2chains\": %.2f, "
           "\"cycles_per_dfma_4chains\": %.2f}\n",
           rate, 2 * rate * 1e-12, p.multiProcessorCount, best, rate / p.multiProcessorCount / (p.clockRate * 1e3),
           p.clockRate / 1000, (double)c1 / ITERS, (double)c2 / (2.0 * ITERS), (double)c4 / (4.0 * ITERS));
    return 0;
}

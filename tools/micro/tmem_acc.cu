// Microbenchmark: per-warp fp64 accumulators resident in tensor memory (TMEM), read-modify-written with
// tcgen05.ld / tcgen05.st.  Question: can 16 warps per SM each RMW 38 columns per ~170 cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_acc tmem_acc.cu && ./tmem_acc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int NACC = 19;            // doubles per lane -> 38 32-bit columns
constexpr int COLS_PER_WARP = 40;
constexpr int WARPS = 16;
constexpr int TMEM_COLS = 256;      // power of two >= (WARPS/4) * COLS_PER_WARP = 160

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// RMW of 8 doubles (16 columns) at column offset c0
__device__ __forceinline__ void rmw8(uint32_t taddr, double w, const double* f) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    wait_ld();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double a = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
        a = fma(w, f[k], a);
        r[2 * k] = (uint32_t)__double2loint(a);
        r[2 * k + 1] = (uint32_t)__double2hiint(a);
    }
    tmem_st16(taddr, r);
}

__global__ void __launch_bounds__(WARPS * 32, 1) tmem_kernel(double* out, int iters, long long* cycles) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * COLS_PER_WARP);
    // zero the accumulators
    {
        uint32_t z[16] = {0};
        tmem_st16(taddr, z);
        tmem_st16(taddr + 16, z);
        uint32_t z8[8] = {0};
        tmem_st8(taddr + 32, z8);
        wait_st();
    }
    double f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 1.0 + k + lane * 0.01;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const double w = 1.0 + 1e-3 * it;
        rmw8(taddr, w, f);        // doubles 0..7
        rmw8(taddr + 16, w, f);   // doubles 8..15
        {                         // doubles 16..19 (we use 19)
            uint32_t r[8];
            tmem_ld8(taddr + 32, r);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double a = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
                a = fma(w, f[k], a);
                r[2 * k] = (uint32_t)__double2loint(a);
                r[2 * k + 1] = (uint32_t)__double2hiint(a);
            }
            tmem_st8(taddr + 32, r);
        }
        wait_st();
    }
    const long long t1 = clock64();
    // read back
    uint32_t r[16];
    tmem_ld16(taddr, r);
    wait_ld();
    out[(blockIdx.x * WARPS + warp) * 32 + lane] = __hiloint2double((int)r[1], (int)r[0]);
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

int main() {
    double* out;
    long long* cyc;
    const int blocks = 148, iters = 20000;
    cudaMalloc(&out, sizeof(double) * blocks * WARPS * 32);
    cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    tmem_kernel<<<blocks, WARPS * 32>>>(out, 100, cyc);
    cudaEventRecord(e0);
    tmem_kernel<<<blocks, WARPS * 32>>>(out, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(err));
        return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    double h[64];
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double expect = 0;   // sum_it (1 + 1e-3 it) * f[0] for lane 0
    for (int it = 0; it < iters; ++it) expect += (1.0 + 1e-3 * it) * 1.0;
    printf("ms %.3f  cycles/iter (one warp view, %d warps/SM) %.1f  lane0 acc %.6f expect %.6f  lane1 %.6f\n", ms, WARPS,
           (double)c / iters, h[0], expect, h[1]);
    return 0;
}

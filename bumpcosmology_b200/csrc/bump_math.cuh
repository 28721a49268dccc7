// Scalar fp64 math for the streaming kernel: lean, branch-free routines with only the special-case handling
// the kernel needs.  Each is accurate to ~1-2 ulp on its stated domain (tests/test_gpu_parity.py::test_device_math
// compares them with CUDA libm through bump_debug_math).  The FP64 pipe (64 lanes/clk/SM on B200) is the
// binding resource of the fp64 path, so these are sized in DFMA-pipe instructions:
//   fexp  10   (libm exp ~18-20 + branches)      frcp  4 + MUFU.RCP64H   (IEEE division ~24)
#pragma once
#include <math.h>

#include "bump_layout.cuh"

namespace bump {

// Polynomial coefficients live in the constant bank: DFMA takes a c[bank][offset] operand directly, whereas a
// literal whose low 32 bits are non-zero costs two UMOV/IMAD.MOV per use (measured: ~75 extra instructions per
// sample, on an issue port that the FP64 stream already half fills).
__constant__ double K_EXP[8] = {
    92.33248261689366,        // [0] 64/ln2
    0.010830424609594047,     // [1] ln2/64 high part (low 26 mantissa bits zero)
    8.66550983900947e-11,     // [2] ln2/64 low part
    6755399441055744.0,       // [3] 1.5 * 2^52
    0.008333333333333333,     // [4] 1/5!
    0.041666666666666664,     // [5] 1/4!
    0.16666666666666666,      // [6] 1/3!
    0.0,
};
__constant__ double K_L1P[4] = {1.0 / 7.0, -1.0 / 6.0, 0.2, 1.0 / 3.0};

// ---- reciprocal of a positive normal double: MUFU.RCP64H seed (~2^-23) + 2 Newton steps
__device__ __forceinline__ double frcp(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- exp(x) for finite x in (-1e5, 700); below about -700 the result saturates at ~1e-304 (callers treat it as
// zero).  x = n (ln2/64) + r, |r| <= ln2/128;  exp(x) = 2^(n>>6) * T[n&63] * (1 + p(r)),  T[j] = 2^(j/64) in shared
// memory (512 bytes; neighbouring lanes carry neighbouring samples, so their arguments mostly share an entry),
// p = degree-5 Taylor polynomial of expm1 (truncation 0.0054^6/6! = 3.5e-17).  The underflow clamp is applied to
// the integer n (one VIMNMX) instead of to x (DSETP + 2 FSEL on the FP64 pipe).  10 FP64-pipe instructions.
__device__ __forceinline__ double fexp(const double x, const double* __restrict__ expt) {
    double kd = fma(x, K_EXP[0], K_EXP[3]);
    const int n = __double2loint(kd);
    kd -= K_EXP[3];
    double r = fma(-kd, K_EXP[1], x);
    r = fma(-kd, K_EXP[2], r);
    double p = K_EXP[4];
    p = fma(p, r, K_EXP[5]);
    p = fma(p, r, K_EXP[6]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    const double T = expt[n & (NEXPT - 1)];
    const double v = fma(T, p, T);
    const int scale = (max(n, -1010 * NEXPT) << 14) & 0xfff00000;   // ((n >> 6) << 20)
    return __hiloint2double(__double2hiint(v) + scale, __double2loint(v));
}

// ---- log(1 + x) for 0 <= x <= 0.0046 (position inside one bin of the log-uniform z grid):
// alternating series to x^7 (remainder 0.0046^8/8 = 2.5e-20)
__device__ __forceinline__ double flog1p_small(const double x) {
    double p = K_L1P[0];
    p = fma(p, x, K_L1P[1]);
    p = fma(p, x, K_L1P[2]);
    p = fma(p, x, -0.25);
    p = fma(p, x, K_L1P[3]);
    p = fma(p, x, -0.5);
    p = fma(p, x, 1.0);
    return p * x;
}

// ---- 1 / (1 + x) for the same range: geometric series to x^7 (remainder 5e-19 relative)
__device__ __forceinline__ double frcp1p_small(const double x) {
    double p = -1.0;
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    return p;
}

}  // namespace bump

// Scalar fp64 math for the streaming kernel: lean, branch-free routines with only the special-case handling
// the kernel needs.  Each is accurate to ~1-2 ulp on its stated domain (tests/test_gpu_parity.py::test_device_math
// compares them with CUDA libm through bump_debug_math).  The FP64 pipe (64 lanes/clk/SM on B200) is the
// binding resource of the fp64 path, so these are sized in DFMA-pipe instructions:
//   fexp  12   (libm exp ~18-20 + branches)      frcp  4 + MUFU.RCP64H   (IEEE division ~24)
#pragma once
#include <math.h>

namespace bump {

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

// ---- exp(x) for finite x <= ~700 (x below -700 is clamped: result ~1e-304, callers treat it as zero).
// x = n (ln2/16) + r, |r| <= ln2/32;  exp(x) = 2^(n>>4) * T[n&15] * (1 + p(r)),  T[j] = 2^(j/16) in shared
// memory (16 doubles = exactly one row of the 32 banks: any access pattern is conflict-free),
// p = degree-7 Taylor polynomial of expm1 (truncation 0.0217^8/8! = 1.2e-18).
__device__ __forceinline__ double fexp(double x, const double* __restrict__ expt) {
    const double INV = 23.083120654223414;            // 16/ln2
    const double C_HI = 0.04332169867120683;          // ln2/16, low 24 mantissa bits zero
    const double C_LO = 1.1378974990650914e-10;
    const double SHIFT = 6755399441055744.0;          // 1.5 * 2^52
    x = fmax(x, -700.0);
    double kd = fma(x, INV, SHIFT);
    const int n = __double2loint(kd);
    kd -= SHIFT;
    double r = fma(-kd, C_HI, x);
    r = fma(-kd, C_LO, r);
    double p = 0.0001984126984126984;                 // 1/7!
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    const double T = expt[n & 15];
    const double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + ((n >> 4) << 20), __double2loint(v));
}

// ---- log(1 + x) for 0 <= x <= 0.0046 (position inside one bin of the log-uniform z grid):
// alternating series to x^7 (remainder 0.0046^8/8 = 2.5e-20)
__device__ __forceinline__ double flog1p_small(const double x) {
    double p = 1.0 / 7.0;
    p = fma(p, x, -1.0 / 6.0);
    p = fma(p, x, 0.2);
    p = fma(p, x, -0.25);
    p = fma(p, x, 1.0 / 3.0);
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

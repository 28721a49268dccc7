// Scalar fp64 math used by the streaming kernel.
//
// BUMP_FAST_MATH=0: CUDA libm (exp/log/log1p and IEEE division) — the first-correct configuration.
// BUMP_FAST_MATH=1: lean table-free/table-light routines (no special-case handling beyond what the kernel
//                   needs: exp(-inf)=0, positive finite arguments for log), accurate to ~1e-15 relative,
//                   validated against libm in tests/test_gpu_math.py.
#pragma once
#include <math.h>

#ifndef BUMP_FAST_MATH
#define BUMP_FAST_MATH 0
#endif

namespace bump {

#if !BUMP_FAST_MATH

__device__ __forceinline__ double fexp(double x) { return exp(x); }
__device__ __forceinline__ double flog(double x) { return log(x); }
__device__ __forceinline__ double frcp(double x) { return 1.0 / x; }
// log(1 + x) for 0 <= x <= ~0.005 (position inside one z-grid bin)
__device__ __forceinline__ double flog1p_small(double x) { return log1p(x); }
// 1 / (1 + x) for the same range
__device__ __forceinline__ double frcp1p_small(double x) { return 1.0 / (1.0 + x); }

#else

// ---- reciprocal: MUFU.RCP64H seed + 2 Newton steps (full double precision for normal inputs)
__device__ __forceinline__ double frcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- exp: x = k ln2 + r, |r| <= ln2/2, degree-11 Taylor-like minimax (coefficients = 1/n!, error < 2e-16 rel
// after Horner on |r| <= 0.3466 with the 12th-order remainder 0.3466^12/12! = 6e-15 ... see tests) .
// Handles x = -inf (-> 0) and large negative x (-> 0) without branches; no overflow handling (callers bound x).
__device__ __forceinline__ double fexp(double x) {
    const double L2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
    double xc = fmax(x, -745.0);
    double kd = fma(xc, L2E, SHIFT);
    int k = __double2loint(kd);
    kd -= SHIFT;
    double r = fma(-kd, LN2_HI, xc);
    r = fma(-kd, LN2_LO, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.0876756987868099e-09);        // 1/12!
    p = fma(p, r, 2.5052108385441719e-08);        // 1/11!
    p = fma(p, r, 2.7557319223985891e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);        // 1/8!
    p = fma(p, r, 1.9841269841269841e-04);        // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);        // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // scale by 2^k in two steps so that k down to -1074 stays correct (results may be subnormal/zero)
    int k1 = k >> 1, k2 = k - k1;
    double s1 = __hiloint2double((k1 + 1023) << 20, 0);
    double s2 = __hiloint2double((k2 + 1023) << 20, 0);
    return (p * s1) * s2;
}

// ---- log for positive finite normal x: x = 2^e m, m in [sqrt(1/2), sqrt(2)); f = (m-1)/(m+1);
// log m = 2 f (1 + f^2/3 + f^4/5 + ...), |f| <= 0.1716, terms to f^22.
__device__ __forceinline__ double flog(double x) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) {  // m >= ~sqrt(2): halve
        hi -= 0x00100000;
        e += 1;
    }
    double m = __hiloint2double(hi, lo);
    double f = (m - 1.0) * frcp(m + 1.0);
    double f2 = f * f;
    double p = 1.0 / 23.0;
    p = fma(p, f2, 1.0 / 21.0);
    p = fma(p, f2, 1.0 / 19.0);
    p = fma(p, f2, 1.0 / 17.0);
    p = fma(p, f2, 1.0 / 15.0);
    p = fma(p, f2, 1.0 / 13.0);
    p = fma(p, f2, 1.0 / 11.0);
    p = fma(p, f2, 1.0 / 9.0);
    p = fma(p, f2, 1.0 / 7.0);
    p = fma(p, f2, 1.0 / 5.0);
    p = fma(p, f2, 1.0 / 3.0);
    p = p * f2;
    double r = fma(2.0 * f, p, 2.0 * f);
    return fma((double)e, 0.69314718055994530942, r);
}

__device__ __forceinline__ double flog1p_small(double x) {
    // |x| <= 0.0046: alternating series to x^7 (remainder 0.0046^8/8 = 2.5e-20)
    double p = 1.0 / 7.0;
    p = fma(p, x, -1.0 / 6.0);
    p = fma(p, x, 0.2);
    p = fma(p, x, -0.25);
    p = fma(p, x, 1.0 / 3.0);
    p = fma(p, x, -0.5);
    p = fma(p, x, 1.0);
    return p * x;
}

__device__ __forceinline__ double frcp1p_small(double x) {
    // 1/(1+x), |x| <= 0.0046: geometric series to x^7 (remainder 5e-19 relative)
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

#endif

}  // namespace bump

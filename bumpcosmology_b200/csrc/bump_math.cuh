// Scalar fp64 math for the streaming kernel: lean, branch-free routines with only the special-case handling
// the kernel needs.  Each is accurate to ~1-2 ulp on its stated domain (tests/test_gpu_parity.py::test_device_math
// compares them with CUDA libm through bump_debug_math).  The FP64 pipe (64 lanes/clk/SM on B200) is the
// binding resource of the fp64 path, so these are sized in DFMA-pipe instructions:
//   fexp  7-8  (libm exp ~18-20 + branches)      frcp  4 + MUFU.RCP64H   (IEEE division ~24)
#pragma once
#include <math.h>

#include "bump_layout.cuh"

namespace bump {

// Constants live in the constant bank: DFMA takes a c[bank][offset] operand directly, whereas a literal whose low
// 32 bits are non-zero costs two UMOV/IMAD.MOV per use (measured: ~75 extra instructions per sample, on an issue
// port that the FP64 stream already fills; profiles/r01_fp64_issue_microbench.txt).
__constant__ double K_EXP[4] = {
    2954.639443740597,         // [0] 2048/ln2
    0.0003384507717577858,     // [1] ln2/2048, correctly rounded
    1.1323470770733885e-20,    // [2] ln2/2048 - [1]   (second step of the wide-range reduction)
    0.16666666666666666,       // [3] 1/3!
};
__constant__ double K_L1P[4] = {1.0 / 7.0, -1.0 / 6.0, 0.2, 1.0 / 3.0};

// ---- shared-memory loads by 32-bit shared-window address + compile-time byte offset.  The table blob is addressed
// as (laundered base register) + index * size with the table's offset folded into the instruction: no generic ->
// shared conversion, and the compiler cannot rematerialise the base (S2R + MOV + LEA) at every use.
template <int OFF>
__device__ __forceinline__ double lds64(const uint32_t a) {
    double v;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ double2 lds128(const uint32_t a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds16(const uint32_t a) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}

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

// ---- exp(x) for finite x in (-1e5, 700).  x = n (ln2/2048) + r, |r| <= ln2/4096;  exp(x) = 2^(n>>11) * T[n&2047]
// * (1 + p(r)),  T[j] = 2^(j/2048) in shared memory (16 KB, correctly rounded on the host), p = degree-3 Taylor
// polynomial of expm1 (truncation (1.7e-4)^4/4! = 3.4e-17).  Below about -700 the result saturates at ~1e-304
// (callers treat it as zero): the clamp is applied to the integer n (one VIMNMX) instead of to x (DSETP + 2 FSEL on
// the FP64 pipe).  `sb` is the shared-window address of the table blob.
//   WIDE = false: one-constant argument reduction, error |x| * 1.1e-16 in r: for the exponents of the mass function
//                 and of the rate (|x| < ~60 wherever the result matters).  7 FP64-pipe instructions.
//   WIDE = true : two-constant reduction, exact over the whole range.  8 FP64-pipe instructions.
template <bool WIDE>
__device__ __forceinline__ double fexp(const double x, const uint32_t sb) {
    constexpr double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: low word of (t + MAGIC) = round(t)
    double kd = fma(x, K_EXP[0], MAGIC);
    const int n = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(-kd, K_EXP[1], x);
    if constexpr (WIDE) r = fma(-kd, K_EXP[2], r);
    double p = fma(r, K_EXP[3], 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    const double T = lds64<OFF_EXPT * 8>(sb + ((uint32_t)(n & (NEXPT - 1)) << 3));
    const double v = fma(T, p, T);
    int hi;   // hi word of v + ((n >> 11) << 20), as ONE integer multiply-add (ptxas otherwise emits shift + add)
    asm("mad.lo.s32 %0, %1, 512, %2;" : "=r"(hi) : "r"(max(n, -1010 * NEXPT) & ~(NEXPT - 1)), "r"(__double2hiint(v)));
    return __hiloint2double(hi, __double2loint(v));
}
static_assert(NEXPT == 2048, "fexp's constants assume a 2048-entry table");

// ---- log(1 + x) for 0 <= x <= 0.00453 (position inside one bin of the log-uniform z grid):
// alternating series to x^6 (remainder 0.00453^7/7 = 5.6e-18, absolute)
__device__ __forceinline__ double flog1p_small(const double x) {
    double p = fma(x, K_L1P[1], 0.2);
    p = fma(p, x, -0.25);
    p = fma(p, x, K_L1P[3]);
    p = fma(p, x, -0.5);
    p = fma(p, x, 1.0);
    return p * x;
}

// ---- 1 / (1 + x) for the same range: geometric series to x^6 (remainder 3.9e-17 relative)
__device__ __forceinline__ double frcp1p_small(const double x) {
    double p = x - 1.0;
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    return p;
}

}  // namespace bump

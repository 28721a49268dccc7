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
constexpr double LN2_HI = 0.69314718055994530942;      // ln 2 correctly rounded
constexpr double LN2_LO = 2.3190468138462996e-17;      // ln 2 - LN2_HI
__constant__ double K_EXP[6] = {
    NEXPT * 1.4426950408889634074,   // [0] NEXPT/ln2
    LN2_HI / NEXPT,                  // [1] ln2/NEXPT (a power-of-two scaling of LN2_HI: still correctly rounded)
    LN2_LO / NEXPT,                  // [2] ln2/NEXPT - [1]   (second step of the wide-range reduction)
    0.16666666666666666,             // [3] 1/3!
    0.041666666666666664,            // [4] 1/4!   (tables shorter than 1024 entries)
    0.0,
};
__constant__ double K_T1P[4] = {1.0 / 7.0, -1.0 / 6.0, 0.2, 1.0 / 3.0};
__constant__ double K_L1P[5] = {0.99999999999999938502, -0.49999999999321161857, 0.33333332134013501958,
                                -0.24999258030344991764, 0.19812434621351388767};
__constant__ double K_R1P[6] = {0.99999999999999999584, -0.99999999999993383492, 0.99999999982954148957,
                                -0.99999983934226992926, 0.99993151458143535823, -0.98652464466049917445};

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

// ---- reciprocal of a positive normal double: MUFU.RCP64H seed (relative error e0 <= 2^-20) + ONE third-order step
// r (1 + e + e^2), e = 1 - x r: remaining error e0^3 <= 2^-60 (+ the roundings of the last FMA), 3 FP64 instructions
// (round 1: two Newton steps, 4).
__device__ __forceinline__ double frcp(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifdef BUMP_RCP_NEWTON2
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#endif
}

// ---- exp(x) for finite x in (-1e5, 700).  x = n (ln2/NEXPT) + r, |r| <= ln2/(2 NEXPT);  exp(x) = 2^(n / NEXPT) *
// T[n mod NEXPT] * (1 + p(r)),  T[j] = 2^(j/NEXPT) in shared memory (correctly rounded on the host), p = Taylor
// polynomial of expm1: degree 3 for NEXPT >= 1024 (truncation r^4/4! = 3.4e-17 at 2048, 5.5e-16 at 1024 entries),
// degree 4 below (3.8e-17 at 256).  Below about -700 the result saturates at ~1e-304 (callers treat it as zero): the
// clamp is applied to the integer n (one VIMNMX) instead of to x (DSETP + 2 FSEL on the FP64 pipe).  `sb` is the
// shared-window address of the table blob, `rep` the byte offset of this lane's copy of the table entries
// ((lane mod EXPT_REPL) * 8; bump_layout.cuh).
//   WIDE = false: one-constant argument reduction, error |x| * 1.1e-16 in r: for the exponents of the mass function
//                 and of the rate (|x| < ~260 wherever the result matters).  7 FP64-pipe instructions (8 with degree 4).
//   WIDE = true : two-constant reduction, exact over the whole range.  One more.
template <bool WIDE>
__device__ __forceinline__ double fexp(const double x, const uint32_t sb, const uint32_t rep = 0u) {
    constexpr double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: low word of (t + MAGIC) = round(t)
    double kd = fma(x, K_EXP[0], MAGIC);
    const int n = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(-kd, K_EXP[1], x);
    if constexpr (WIDE) r = fma(-kd, K_EXP[2], r);
    double p;
    if constexpr (NEXPT >= 1024) {
        p = fma(r, K_EXP[3], 0.5);
    } else {
        p = fma(r, K_EXP[4], K_EXP[3]);
        p = fma(p, r, 0.5);
    }
    p = fma(p, r, 1.0);
    p *= r;
    constexpr int SH = 3 + BUMP_EXPT_REPL_LOG2;   // bytes per table row
    uint32_t off;
    if constexpr (EXPT_REPL > 1) {   // (row offset & mask) | this lane's copy, as ONE three-input logic instruction
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(off) : "r"((uint32_t)n << SH), "n"((NEXPT - 1) << SH), "r"(rep));
    } else {
        off = ((uint32_t)n << SH) & (uint32_t)((NEXPT - 1) << SH);
    }
    const double T = lds64<OFF_EXPT * 8>(sb + off);
    const double v = fma(T, p, T);
    int hi;   // hi word of v + ((n >> log2 NEXPT) << 20), as ONE integer multiply-add (ptxas otherwise emits shift + add)
    asm("mad.lo.s32 %0, %1, %3, %2;" : "=r"(hi) : "r"(max(n, -1010 * NEXPT) & ~(NEXPT - 1)), "r"(__double2hiint(v)),
        "n"((1 << 20) / NEXPT));
    return __hiloint2double(hi, __double2loint(v));
}

// ---- log(1 + x) and 1 / (1 + x) for 0 <= x <= 0.00453 (position inside one bin of the log-uniform z grid).
// Default: Taylor series to x^6 (remainders 5.6e-18 absolute / 3.9e-17 relative) - one instruction more than the
// near-minimax polynomials above, but 7 of its 11 coefficients (1, -1, -1/2, -1/4) are 32-bit immediates, while every
// other fp64 constant costs a load into a register inside the (divergent) sample loop: measured, the series wins.
#ifdef BUMP_POLY_MINIMAX
__device__ __forceinline__ double flog1p_small(const double x) {
    double p = fma(x, K_L1P[4], K_L1P[3]);
    p = fma(p, x, K_L1P[2]);
    p = fma(p, x, K_L1P[1]);
    p = fma(p, x, K_L1P[0]);
    return p * x;
}
__device__ __forceinline__ double frcp1p_small(const double x) {
    double p = fma(x, K_R1P[5], K_R1P[4]);
    p = fma(p, x, K_R1P[3]);
    p = fma(p, x, K_R1P[2]);
    p = fma(p, x, K_R1P[1]);
    p = fma(p, x, K_R1P[0]);
    return p;
}
#else
__device__ __forceinline__ double flog1p_small(const double x) {
    double p = fma(x, K_T1P[1], 0.2);
    p = fma(p, x, -0.25);
    p = fma(p, x, K_T1P[3]);
    p = fma(p, x, -0.5);
    p = fma(p, x, 1.0);
    return p * x;
}
__device__ __forceinline__ double frcp1p_small(const double x) {
    double p = x - 1.0;
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    p = fma(p, x, -1.0);
    p = fma(p, x, 1.0);
    return p;
}
#endif

}  // namespace bump

// The streaming kernel: one pass over the group-blocked sample / injection columns producing, per (warp, event) record,
// the shifted sums
//   S = sum e^{w-m},  S2 = sum e^{2(w-m)}  and the 17 gradient features  sum e^{w-m} f_k.
//
// Replaces intensity_models.py:378-381 (events) and :385-388 (injections) — z_of_dL, detector->source masses,
// LogDNDMDQDV.__call__ (:202-210), LogDNDM.__call__ (:140-151), log_smooth_turnon (:45-54), LogDNDV.__call__
// (:170-173), the Jacobian terms — plus the inner part of the logsumexp reductions (:382,389,392,401) and the
// reverse pass of all of it, in forward mode.
//
// Arithmetic is organised in LINEAR space: the weight of a sample is the product
//   p = e^{w-m} = (e^{P(m1)} + e^{Q(m1)}) (e^{P(m2)} + e^{Q(m2)}) * dVc/dz / (d dL/dz) / (1+r) * e^{lin-m}
// where lin collects every term that is linear in precomputed logs, so a sample costs 8 exp, 4 reciprocals
// and NO logarithm; softmax-weighted gradient terms are products of the same factors (no divisions).
// The shift m is per thread: the `lin` of its first finite-weight sample, raised only when a later sample
// exceeds it by e^RESCALE_GAP (fp64 has the range to carry everything else); lanes, records and ranks are merged
// with the usual (max, rescale) rule, so the result equals the reference's max-shifted logsumexp.
//
// What bounds it (profiles/r01_fp64_issue_microbench.txt): the issue port.  On B200 an FP64
// instruction occupies its sub-partition for max(2, distinct vector-register operands) cycles and nothing co-issues
// beside it, so the code below is written to minimise instruction COUNT: constants come from the constant bank
// (K_SC, K_EXP), table addresses are base register + immediate (lds64/lds128<OFF>), loads carry no predicate (whole
// groups, blocked layout), and the math is folded into single FMAs wherever two constants do not collide.
#pragma once
#include <cuda_runtime.h>

#include "bump_layout.cuh"
#include "bump_math.cuh"

namespace bump {

#ifndef BUMP_STREAM_THREADS
#define BUMP_STREAM_THREADS 384
#endif
constexpr int STREAM_THREADS = BUMP_STREAM_THREADS;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;
// The blob is staged at its natural offsets, minus its first NSCAL doubles: the streaming kernel takes the scalars
// from the constant bank, so their 512 bytes of shared memory hold the mbarrier instead.
__host__ __device__ constexpr int stream_smem_bytes(const bool wa, const bool fixed) {
    return blob_doubles(wa, fixed) * 8;
}
constexpr double ZERO_WEIGHT_SHIFT = -5.0e4;   // exponent shift of a sample without weight: exp saturates (~1e-304)
constexpr double RESCALE_GAP = 200.0;   // p <= e^200 * O(e^50): p^2 stays far below DBL_MAX

// The theta-dependent scalars of the current evaluation, copied device-to-device from the table blob right before
// the launch: FP64 instructions take c[bank][offset] operands directly, so a scalar costs neither a shared-memory
// load nor a register.  One evaluation at a time per slot (bump_lib.cu chains launches through an event).
// NSLOT copies, one per constant SLOT: a context is bound to a slot at creation and the kernels are instantiated per
// slot, so evaluations of up to NSLOT contexts (parallel NUTS chains on a small catalog) overlap on one device.
constexpr int NSLOT = 4;
__constant__ double K_SC4[NSLOT][NSCAL];
#ifndef BUMP_SCALARS_FROM_BLOB
#define K_SC K_SC4[SLOT]   // inside templates with an `int SLOT` parameter
#define USC_PARAM
#define USC_ARG
#else
#define K_SC usc           // warp-uniform copies of the blob's scalars, made at kernel start (see stream_kernel)
#define USC_PARAM const double (&usc)[NSCAL],
#define USC_ARG usc,
#endif

// ---- TMA bulk copy (global -> shared) of the table blob, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BLOB_BYTES, int SKIP_BYTES>
__device__ __forceinline__ void stage_tables(double* s_blob, uint64_t* mbar, const double* __restrict__ g_blob) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                     "r"((uint32_t)(BLOB_BYTES - SKIP_BYTES))
                     : "memory");
        constexpr int CHUNK = 32768;
        for (int off = SKIP_BYTES; off < BLOB_BYTES; off += CHUNK) {
            const int n = (BLOB_BYTES - off < CHUNK) ? (BLOB_BYTES - off) : CHUNK;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(reinterpret_cast<char*>(s_blob) + off)),
                "l"(reinterpret_cast<const char*>(g_blob) + off), "r"((uint32_t)n), "r"(smem_u32(mbar))
                : "memory");
        }
    }
    // every thread waits for phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar))
            : "memory");
    }
}

template <int V>
struct IC {   // integral constant (table ids as template arguments of generic lambdas)
    static constexpr int value = V;
};

struct ThreadAcc {
    double m;           // shift
    double a[NACC];     // S, S2, features (BUMP_TMEM_ACC: live only between a chunk's load and store inside a sample)
    int nvalid;
    uint32_t taddr;     // BUMP_TMEM_ACC: this warp's accumulator columns in tensor memory
};

#ifdef BUMP_TMEM_ACC
// ---- The 19 fp64 sums of a lane in TENSOR MEMORY instead of registers (no tensor-core math involved).  They cost 38
// registers for the whole kernel and cap the CTA at 12 warps; at 128 registers 16 warps fit, and the kernel - bound
// by the latencies of the FP64 pipe and of shared memory, no longer by issue slots - runs as fast as its warps hide
// them (10 -> 12 warps per SM was +16 %).  Each warp owns 40 columns x its 32 lanes of the SM's tensor memory, as two
// chunks of 10 doubles that a sample loads right before it updates them and stores right after:
//   chunk R (columns 0..19) : S, S2, F_CZ, F_OM, F_W, F_WA, F_BETA, F_L, F_SIG, F_SIGL
//   chunk M (columns 20..39): F_SQ, F_C, F_T, F_GEO, F_PA .. F_PSIGMA, (spare)
constexpr int TMEM_COLS_PER_WARP = 40;
constexpr int TMEM_ALLOC_COLS = 256;       // power of two >= (STREAM_WARPS / 4) * TMEM_COLS_PER_WARP
static_assert((STREAM_WARPS + 3) / 4 * TMEM_COLS_PER_WARP <= TMEM_ALLOC_COLS, "tensor-memory columns");
__device__ constexpr int ACC_CHUNK_R[10] = {0, 1, 2 + F_CZ, 2 + F_OM, 2 + F_W, 2 + F_WA, 2 + F_BETA, 2 + F_L, 2 + F_SIG, 2 + F_SIGL};
__device__ constexpr int ACC_CHUNK_M[10] = {2 + F_SQ, 2 + F_C, 2 + F_T, 2 + F_GEO, 2 + F_PA, 2 + F_PB, 2 + F_PMPISN,
                                            2 + F_PMBHMAX, 2 + F_PSIGMA, -1};

__device__ __forceinline__ void tmem_ld20(const uint32_t taddr, uint32_t (&r)[20]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19])
                 : "r"(taddr + 16));
}
__device__ __forceinline__ void tmem_st20(const uint32_t taddr, const uint32_t (&r)[20]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr + 16), "r"(r[16]), "r"(r[17]),
                 "r"(r[18]), "r"(r[19]));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Warp-collective.  MASS = the chunk of the mass-function sums, else the chunk with S, S2 and the rest.
template <bool MASS>
__device__ __forceinline__ void acc_load(ThreadAcc& A) {
    uint32_t r[20];
    tmem_wait_st();   // the store of the previous sample to these columns
    tmem_ld20(A.taddr + (MASS ? 20 : 0), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = MASS ? ACC_CHUNK_M[i] : ACC_CHUNK_R[i];
        if (k >= 0) A.a[k] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
    }
}
template <bool MASS>
__device__ __forceinline__ void acc_store(const ThreadAcc& A) {
    uint32_t r[20];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = MASS ? ACC_CHUNK_M[i] : ACC_CHUNK_R[i];
        r[2 * i] = k >= 0 ? (uint32_t)__double2loint(A.a[k]) : 0u;
        r[2 * i + 1] = k >= 0 ? (uint32_t)__double2hiint(A.a[k]) : 0u;
    }
    tmem_st20(A.taddr + (MASS ? 20 : 0), r);
}
#define ACC_LOAD_MASS(A) acc_load<true>(A)
#define ACC_STORE_MASS(A) acc_store<true>(A)
#define ACC_LOAD_REST(A) acc_load<false>(A)
#define ACC_STORE_REST(A) acc_store<false>(A)
#else
#define ACC_LOAD_MASS(A)
#define ACC_STORE_MASS(A)
#define ACC_LOAD_REST(A)
#define ACC_STORE_REST(A)
#endif

__device__ __forceinline__ void acc_init(ThreadAcc& A) {
    A.m = -INFINITY;
    A.nvalid = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) A.a[k] = 0.0;
#ifdef BUMP_TMEM_ACC
    acc_store<true>(A);
    acc_store<false>(A);
#endif
}

// Raise the shift of the accumulators to `lin` (rare: the first finite-weight sample of a record, or one that exceeds
// the shift by e^RESCALE_GAP).  With the sums in tensor memory this is a warp-collective read-modify-write: lanes that
// do not need it scale by 1.
template <class Exp>
__device__ __forceinline__ void acc_rescale(ThreadAcc& A, const bool need, const double lin, Exp&& exp_wide) {
#ifdef BUMP_TMEM_ACC
    if (!__any_sync(0xffffffffu, need)) return;
    const double s = !need ? 1.0 : ((A.m == -INFINITY) ? 0.0 : exp_wide(A.m - lin));
    acc_load<true>(A);
    acc_load<false>(A);
#else
    if (!need) return;
    const double s = (A.m == -INFINITY) ? 0.0 : exp_wide(A.m - lin);
#endif
    A.a[0] *= s;
    A.a[1] *= s * s;
#pragma unroll
    for (int k = 2; k < NACC; ++k) A.a[k] *= s;
    if (need) A.m = lin;
#ifdef BUMP_TMEM_ACC
    acc_store<true>(A);
    acc_store<false>(A);
#endif
}

// One mass-function evaluation in linear space: e^{A0(m)} = EP + EQ with EP = e^{PISN(m)} (:110-111,144-145),
// EQ = e^{-c log(m/mbhmax) + log_pl_norm + turnon(m)} (:147, :45-54).
struct MassEval {
    double EP, EQ;   // the two terms of the logaddexp, in linear space
    double lrel;     // log(m / mbhmax)
    double sgm;      // m * e/(1+e) : m times the turn-on's logistic weight
    double gy;       // G_{b+1} - G_b: dP/dm inside the bin, per grid step
    double pos;      // (m - 3) / (grid step): position on the mbh grid
    double m, u;
    uint32_t b;      // shared-window address of the bin's records (blob base + 16 b)
};

constexpr int MASS_BYTES = OFF_MASS * 8;   // byte offsets of the tables inside the blob
constexpr int COS_BYTES = OFF_COS * 8;
constexpr int CTAN_BYTES = OFF_CTAN * 8;
constexpr int SRCH_BYTES = OFF_SRCH * 8;

// `sb` = shared-window address of the table blob.  SHIFTED: both exponents carry the extra term `d` (the sample's
// lin - shift for the evaluation at m1, see eval_sample), i.e. the result is e^d (EP + EQ) at no extra exponential.
template <int SLOT, bool SHIFTED>
__device__ __forceinline__ void mass_eval(USC_PARAM const double m, const double lm, const double d, const uint32_t sb,
                                          const uint32_t rep, MassEval& o) {
    // -(m - M)/(0.05 M) = 20 - m/(0.05 M): one FMA
    const double e = fexp<false>(fma(m, -K_SC[S_INV_DM], 1.0 / TURNON_WIDTH), sb, rep);
    const double s1 = frcp(1.0 + e);
    o.sgm = (e * s1) * m;
    o.lrel = lm - K_SC[S_LOG_M];
    const double cq = SHIFTED ? K_SC[S_LOG_C2] + d : K_SC[S_LOG_C2];
    o.EQ = fexp<false>(fma(o.lrel, -K_SC[S_C], cq), sb, rep) * s1;   // 2 e^{lpn} (m/M)^-c / (1 + e)
    const double pos = fma(m, K_SC[S_INV_DMBH], K_SC[S_POS0]);
    o.pos = pos;
    int b = __double2int_rd(pos);
    b = min(max(b, 0), NM - 1);                   // NM-1: the beyond-the-grid record (log dN = -5e4, slope 0; :145)
    o.u = pos - (double)b;
    o.b = sb + 16u * (uint32_t)b;
    const double2 g = lds128<MASS_BYTES + MR_G * NM * 16>(o.b);
    o.EP = fexp<false>(fma(o.u, g.y, SHIFTED ? g.x + d : g.x), sb, rep);   // m <= 3 cannot happen once m >= 5
    o.gy = g.y;
    o.m = m;
}

// Feature contributions of one mass evaluation, weighted by wp = (weight of the sample) / (EP + EQ).
template <int K>
__device__ __forceinline__ void mass_tangents(const MassEval& o, const double wP, double* __restrict__ a) {
    if constexpr (K < 5) {
        const double2 t = lds128<MASS_BYTES + (MR_GA + K) * NM * 16>(o.b);
        a[2 + F_PA + K] = fma(wP, fma(o.u, t.y, t.x), a[2 + F_PA + K]);
        mass_tangents<K + 1>(o, wP, a);
    }
}

template <int SLOT>
__device__ __forceinline__ double mass_features(USC_PARAM const MassEval& o, const double wp, double* __restrict__ a) {
    const double wQ = wp * o.EQ, wP = wp * o.EP;
    a[2 + F_SQ] += wQ;
    a[2 + F_C] = fma(wQ, o.lrel, a[2 + F_C]);
    const double wQs = wQ * o.sgm;                         // wQ * m * logistic
    a[2 + F_T] += wQs;
    const double wPg = wP * o.gy;
    const double geo = wPg * o.pos;                 // wP (dP/dm) (m - 3) per grid step
    a[2 + F_GEO] += geo;
    mass_tangents<0>(o, wP, a);
    // weight * m dA0/dm = wP slope m + wQ (m dT/dm - c), with slope * m = gy (pos + 3 / step): pos0 = -3 / step
    return fma(wPg, -K_SC[S_POS0], geo) + fma(wQs, K_SC[S_INV_DM], -K_SC[S_C] * wQ);
}

// Fixed-cosmology variant (the reference's `pop_model`, intensity_models.py:313-355): the sample carries source-frame
// (m1, q) and log1p(z) directly, `lpd` = log pdraw - log dVdzdt(z) was folded at upload, and there is no d_L
// inversion, no Jacobian and no cosmological gradient.
template <int SLOT, bool WA, class Mid>
__device__ __forceinline__ void eval_sample_fixed(USC_PARAM const double L, double m1, const double q, double lm1,
                                                  const double lq, const double l1q, const double lpd,
                                                  const uint32_t sb, const uint32_t rep, ThreadAcc& A, Mid&& mid) {
    double m2 = q * m1;
    double lm2 = lm1 + lq;
    const bool valid = (m1 >= MBH_MIN) && (m2 >= MBH_MIN);   // :149
    const double pair = lm1 + l1q;
    const double lin = fma(K_SC[S_BETA], pair, lm1) + fma(K_SC[S_LAM], L, -lpd);   // :332 (no (1+z)^-2 Jacobian here)
    mid();
    acc_rescale(A, valid && lin - A.m > RESCALE_GAP, lin, [&](const double x) { return fexp<true>(x, sb, rep); });
    const double d = valid ? lin - A.m : ZERO_WEIGHT_SHIFT;   // e^d is folded into the two exponentials at m1
    A.nvalid += valid ? 1 : 0;
    const double r = fexp<false>(K_SC[S_KAPPA] * (L - K_SC[S_LOPZP]), sb, rep);
    const double sr = frcp(1.0 + r);
    MassEval M1, M2;
    mass_eval<SLOT, true>(USC_ARG m1, lm1, d, sb, rep, M1);
    mass_eval<SLOT, false>(USC_ARG m2, lm2, 0.0, sb, rep, M2);
    const double sum1 = M1.EP + M1.EQ, sum2 = M2.EP + M2.EQ;
    const double base = sr;
    const double p = (sum1 * sum2) * base;
    ACC_LOAD_MASS(A);
    mass_features<SLOT>(USC_ARG M1, sum2 * base, A.a);
    mass_features<SLOT>(USC_ARG M2, sum1 * base, A.a);
    ACC_STORE_MASS(A);
    ACC_LOAD_REST(A);
    A.a[0] += p;
    A.a[1] = fma(p, p, A.a[1]);
    const double psig = p * (r * sr);
    A.a[2 + F_BETA] = fma(p, pair, A.a[2 + F_BETA]);
    A.a[2 + F_L] = fma(p, L, A.a[2 + F_L]);
    A.a[2 + F_SIG] += psig;
    A.a[2 + F_SIGL] = fma(psig, L, A.a[2 + F_SIGL]);
    ACC_STORE_REST(A);
}

// `mid()` runs once every input of the sample has been consumed (used by the caller to issue the next loads there:
// issuing them earlier makes the first use of THIS sample's inputs wait on a scoreboard slot shared with the fresh
// loads, i.e. on a full L2 round trip).
template <int SLOT, bool WA, class Mid>
__device__ __forceinline__ void eval_sample(USC_PARAM const double x, const double m1d, const double q, const double lm,
                                            const double lq, const double l1q, const double lpd,
                                            const uint32_t sb, const uint32_t rep, ThreadAcc& A, Mid&& mid) {
    // ---- z_of_dL: b = clip(searchsorted(dl, x, side='right'), 1, n-1) - 1   (:272-273, jnp.interp)
    // The bucket table, keyed by the top bits of x, gives a lower bound b0 of the bin.  A bucket (1/256 octave) is
    // narrower than any bin of the d_L grid (log dl_{k+1} - log dl_k >= ZSTEP = 0.0045 > log(1 + 1/256)), so the
    // bin is b0 or b0 + 1: one comparison with knot b0 + 1, no loop.  (the prologue flags the evaluation as bad
    // if theta is so extreme that the ends of the bucket range break this: roughly h > 7 or h < 0.11; the prior is 0.35 .. 1.4.)
    int j = (__double2hiint(x) >> (20 - SRCH_MBITS)) - (SRCH_EXP_LO << SRCH_MBITS);
    j = min(max(j, 0), SRCH_N - 1);
    const uint32_t b0 = lds16<SRCH_BYTES>(sb + 2u * (uint32_t)j);
    const double knot = lds64<COS_BYTES + CR_DL * NZ * 16 + 16>(sb + 16u * b0);   // dl[b0 + 1]
    // bin NZ-1 is the record BEYOND the table (x >= the last knot): the last knot's values with zero slopes and zero
    // 1/width, i.e. t = 0, idl = 0 - jnp.interp clamps to fp[-1] and no gradient flows to x or xp - with no select
    // here.  (The w0-wa kernel's knot-value tangent tables have no such slot: it clamps to the last bin and selects.)
    const uint32_t b = min(b0 + (x >= knot ? 1u : 0u), (uint32_t)(WA ? NZ - 2 : NZ - 1));
    const uint32_t ab = sb + 16u * b;   // the bin's pair records
    const uint32_t at = sb + 8u * b;    // the bin's tangent-table knots (w0-wa mode)
    const double2 rdl = lds128<COS_BYTES + CR_DL * NZ * 16>(ab);
    double t = (x - rdl.x) * rdl.y;
    double idl = rdl.y;
    if constexpr (WA) {
        const bool beyond = x > K_SC[S_DL_LAST];
        idl = beyond ? 0.0 : rdl.y;
        t = beyond ? 1.0 : t;
    }
    // ---- position inside the z bin: 1+z = (1+z_b)(1 + t eps)
    const double2 rz = lds128<COS_BYTES + CR_Z * NZ * 16>(ab);
    const double zeps = K_SC[S_ZEPS];
    const double te = t * zeps;
    const double u1 = frcp1p_small(te);
    const double ropz = rz.x * u1;                  // 1/(1+z)
    const double L = rz.y + flog1p_small(te);       // log1p(z)
    // ---- source-frame masses (:379, :207)
    double m1 = m1d * ropz;
    double m2 = q * m1;
    double lm1 = lm - L;
    double lm2 = lm1 + lq;
    // ---- dVC/dz and d(dL)/dz lerps at z (:264-268), same bin, same t
    const double2 rvc = lds128<COS_BYTES + CR_DVC * NZ * 16>(ab);
    const double2 rdd = lds128<COS_BYTES + CR_DDL * NZ * 16>(ab);
    const double dvc = fma(t, rvc.y, rvc.x);
    const double ddl = fma(t, rdd.y, rdd.x);
    // Zero weight (:149; log(dVc/dz = 0) = -inf).  Every intermediate of such a sample stays finite without any
    // substitution: samples that can never be valid (m2_det < mbh_min, i.e. m2 < 5 at every redshift) are replaced by
    // sentinels at upload, so here m1 >= m2 >= 5/101 and all exponents below are bounded; the weight is zeroed once.
    // (dVc/dz = 0, i.e. log(dVc/dz) = -inf in the reference, needs no test in linear space: the weight is a product
    // with dvc)
    const bool valid = (m1 >= MBH_MIN) && (m2 >= MBH_MIN);
    const double iddl = frcp(ddl);
    // ---- everything that is linear in precomputed logs: beta log(m1+m2) + log m1 + (lam-2) log1p(z) - log pdraw
    const double pair = lm1 + l1q;                  // log(m1+m2); the -beta log(60) is in the constant
    const double lin = fma(K_SC[S_BETA], pair, lm1) + fma(K_SC[S_LAM2], L, -lpd);
    mid();
    // e^{lin - shift} is not computed on its own: d is added to both exponents of the mass function at m1 (7 instead of
    // 8 exponentials per sample; |d| <= RESCALE_GAP keeps the one-constant argument reduction at ~3e-14 relative)
    // A sample without weight gets d = -5e4: both exponentials at m1 then return exp's saturation value (~1e-304
    // relative to any real weight: below every rounding of the sums) - one select instead of zeroing the weight as well.
    double d = valid ? lin - A.m : ZERO_WEIGHT_SHIFT;
    {
        const bool need = d > RESCALE_GAP;          // also the first finite sample (shift = -inf: d = +inf)
        acc_rescale(A, need, lin, [&](const double x) { return fexp<true>(x, sb, rep); });
        if (need) d = 0.0;
    }
    A.nvalid += valid ? 1 : 0;
    // ---- merger-rate density (:173): (1+z)^lam / (1 + r),  r = ((1+z)/(1+zp))^kappa
    const double kappa = K_SC[S_KAPPA];
    const double r = fexp<false>(kappa * (L - K_SC[S_LOPZP]), sb, rep);
    const double sr = frcp(1.0 + r);
    const double sig = r * sr;
    // ---- mass function at both masses
    MassEval M1, M2;
    mass_eval<SLOT, true>(USC_ARG m1, lm1, d, sb, rep, M1);
    mass_eval<SLOT, false>(USC_ARG m2, lm2, 0.0, sb, rep, M2);
    const double sum1 = M1.EP + M1.EQ, sum2 = M2.EP + M2.EQ;   // sum1 carries e^{lin - shift}
    // ---- the weight and its partial products
    const double base = sr * iddl;                  // everything but the masses and dVc/dz
    const double p0 = (sum1 * sum2) * base;         // weight / (dVc/dz)
    const double p = p0 * dvc;                      // e^{w - m}   (:381 / :388)
    const double bv = base * dvc;
    ACC_LOAD_MASS(A);
    const double md = mass_features<SLOT>(USC_ARG M1, sum2 * bv, A.a) + mass_features<SLOT>(USC_ARG M2, sum1 * bv, A.a);
    ACC_STORE_MASS(A);
    ACC_LOAD_REST(A);
    A.a[0] += p;
    A.a[1] = fma(p, p, A.a[1]);
    // ---- d w / d t at fixed tables (times p), then the cosmological tangents
    const double lt = zeps * u1;                    // d log1p(z) / dt
    const double psig = p * sig;
    const double pWt = fma(lt, fma(p, K_SC[S_RATE0], -fma(kappa, psig, md)), fma(rvc.y, p0, -(rdd.y * iddl) * p));
    const double pWx = pWt * idl;                   // -pWx * (d dl-table/d theta)(t) = p (dw/dt)(dt/dtheta)
    A.a[2 + F_CZ] = fma(pWx, x, A.a[2 + F_CZ]);
    const double pid = p * iddl;
    auto tangent = [&](auto tdl, auto tdvc, auto tddl, double& acc) {
        if constexpr (WA) {   // knot values
            constexpr int ODL = CTAN_BYTES + decltype(tdl)::value * NZ * 8, ODVC = CTAN_BYTES + decltype(tdvc)::value * NZ * 8,
                          ODDL = CTAN_BYTES + decltype(tddl)::value * NZ * 8;
            const double a0 = lds64<ODL>(at), a1 = lds64<ODL + 8>(at);
            const double v0 = lds64<ODVC>(at), v1 = lds64<ODVC + 8>(at);
            const double d0 = lds64<ODDL>(at), d1 = lds64<ODDL + 8>(at);
            acc = fma(-pWx, fma(t, a1 - a0, a0), fma(p0, fma(t, v1 - v0, v0), fma(-pid, fma(t, d1 - d0, d0), acc)));
        } else {              // per-bin pairs {t_b, t_{b+1} - t_b}
            const double2 a = lds128<CTAN_BYTES + decltype(tdl)::value * NZ * 16>(ab);
            const double2 v = lds128<CTAN_BYTES + decltype(tdvc)::value * NZ * 16>(ab);
            const double2 d = lds128<CTAN_BYTES + decltype(tddl)::value * NZ * 16>(ab);
            acc = fma(-pWx, fma(t, a.y, a.x), fma(p0, fma(t, v.y, v.x), fma(-pid, fma(t, d.y, d.x), acc)));
        }
    };
    tangent(IC<CT_DL_OM>{}, IC<CT_DVC_OM>{}, IC<CT_DDL_OM>{}, A.a[2 + F_OM]);
    tangent(IC<CT_DL_W>{}, IC<CT_DVC_W>{}, IC<CT_DDL_W>{}, A.a[2 + F_W]);
    if constexpr (WA) tangent(IC<CT_DL_WA>{}, IC<CT_DVC_WA>{}, IC<CT_DDL_WA>{}, A.a[2 + F_WA]);
    A.a[2 + F_BETA] = fma(p, pair, A.a[2 + F_BETA]);
    A.a[2 + F_L] = fma(p, L, A.a[2 + F_L]);
    A.a[2 + F_SIG] += psig;
    A.a[2 + F_SIGL] = fma(psig, L, A.a[2 + F_SIGL]);
    ACC_STORE_REST(A);
}

// Warp-wide merge of the per-lane accumulators (fixed shuffle tree: deterministic) -> one record.
__device__ __forceinline__ void warp_flush(ThreadAcc& A, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
#ifdef BUMP_TMEM_ACC
    acc_load<true>(A);
    acc_load<false>(A);
#endif
    double mx = A.m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const double sc = (A.m == -INFINITY) ? 0.0 : exp(A.m - mx);
    A.a[0] *= sc;
    A.a[1] *= sc * sc;
#pragma unroll
    for (int k = 2; k < NACC; ++k) A.a[k] *= sc;
    double nv = (double)A.nvalid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    double mine = 0.0;   // lane k keeps the k-th sum
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        double v = A.a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == k) mine = v;
    }
    if (lane < NACC) out[1 + lane] = mine;
    if (lane == NACC) out[1 + NACC] = nv;
    if (lane == 31) out[0] = mx;
}

// Persistent, warp-autonomous streaming kernel.  After the table blob has been staged into shared memory (TMA
// bulk copy, one mbarrier) no warp ever synchronises with another: warp w walks its range of 64-sample groups
// [w*gpw, (w+1)*gpw), lane l evaluating samples l and 32+l of each group (two coalesced 64-bit loads per column),
// and flushes a record whenever the event changes.  Every event (and the injection set) is padded to whole groups
// with zero-weight sentinel samples at upload, so the loads carry no predicate.
#ifdef BUMP_STREAM_MAXREG
#define BUMP_STREAM_BOUNDS __maxnreg__(BUMP_STREAM_MAXREG)
#else
#define BUMP_STREAM_BOUNDS __launch_bounds__(STREAM_THREADS, 1)
#endif
template <bool WA, bool FIXED, int SLOT>
__global__ void BUMP_STREAM_BOUNDS
stream_kernel(const Columns cols, const Work wk, const int* __restrict__ rec_off,
              const double* __restrict__ g_blob, double* __restrict__ part, unsigned long long* __restrict__ tl) {
    pdl_launch_dependents<PDL_EPILOGUE && BUMP_PDL_EPI_TRIGGER == 2>();   // (measurement option: at kernel start)
    timeline_begin(tl, TL_STREAM);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int BLOB_BYTES = blob_doubles(WA, FIXED) * 8;   // the mode's share of the blob (bump_layout.cuh)
    double* s_blob = reinterpret_cast<double*>(smem_raw);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);   // inside the (unused) scalar block
#ifdef BUMP_TMEM_ACC
    // tensor memory for the accumulators: one warp allocates, everybody reads the base address after the barrier
    // inside stage_tables (the address lands in the unused scalar block, behind the mbarrier)
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(smem_raw + 16);
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base)),
                     "n"(TMEM_ALLOC_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
#endif

    // This grid may have been scheduled while the prologue was still running: nothing theta-dependent (the blob, the
    // constant-bank slot) is touched before this returns.
    pdl_wait<PDL_STREAM>();
#ifdef BUMP_SCALARS_FROM_BLOB
    // Build option BUMP_SCALARS_FROM_BLOB: no constant bank at all (no constant-bank slots shared between contexts).
    // Lane l fetches scal[l] and scal[32 + l] from the blob in global memory (the loads
    // fly while the tables are staged), and every scalar the loop needs is broadcast from its lane: a shuffle from a
    // fixed lane is a value the compiler knows to be warp-uniform.  ptxas promotes only three such values to uniform
    // registers, though (it hoists a dozen constant-bank loads into them): measured on B200 the streaming kernel is 4 %
    // slower this way (0.0190 against 0.0183 ns per sample), and the constant bank costs nothing per evaluation (the
    // prologue writes the slot itself), so it stays the default.
    const double sc_lo = __ldcg(g_blob + OFF_SCAL + (threadIdx.x & 31));
    const double sc_hi = __ldcg(g_blob + OFF_SCAL + 32 + (threadIdx.x & 31));
#endif
    stage_tables<BLOB_BYTES, NSCAL * 8>(s_blob, mbar, g_blob);
    timeline_begin(tl, TL_STREAM_STAGED);
    timeline_end(tl, TL_STREAM_STAGED);
#ifdef BUMP_SCALARS_FROM_BLOB
    double usc[NSCAL];
    {
        constexpr int used[] = {S_C, S_LOG_M, S_INV_DM, S_TOP, S_INV_DMBH, S_BETA, S_LAM, S_KAPPA, S_LOPZP, S_DL_LAST,
                                S_ZEPS, S_LAM2, S_RATE0, S_LOG_C2, S_POS0};
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) usc[k] = 0.0;
#pragma unroll
        for (int j = 0; j < (int)(sizeof(used) / sizeof(int)); ++j)
            usc[used[j]] = __shfl_sync(0xffffffffu, used[j] < 32 ? sc_lo : sc_hi, used[j] & 31);
    }
#endif
    // shared-window address of the blob, laundered so that it lives in one register for the whole kernel (the
    // compiler otherwise rematerialises it - S2R, MOV, LEA - in front of every table access)
    uint32_t sb = smem_u32(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb));

#ifdef BUMP_TMEM_ACC
    asm volatile("tcgen05.fence::after_thread_sync;");
    // a warp reaches only its own quarter of the 128 tensor-memory lanes (warp id mod 4); the warps of one quarter get
    // disjoint column ranges
    const uint32_t tmem_addr = *tmem_base + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) +
                               (uint32_t)((threadIdx.x >> 7) * TMEM_COLS_PER_WARP);
#endif
    // everything one warp does after the staging (it may return early; the kernel frame below owns the final barrier
    // of the tensor-memory variant)
    auto warp_work = [&]() {
    const int lane = threadIdx.x & 31;
    const uint32_t rep = (uint32_t)(lane & (EXPT_REPL - 1)) << 3;   // this lane's copy of the exp-table entries
    uint64_t l2_stream_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_stream_policy));
    // The warp index through a shuffle from lane 0: the same value, but one the compiler KNOWS to be warp-uniform.
    // Everything derived from it (the warp's group range, the loop trip count, event boundaries) is then uniform too,
    // the sample loop is convergent code, and ptxas keeps the theta-dependent scalars in UNIFORM registers, which FP64
    // instructions read as a third operand for free.  With `threadIdx.x >> 5` the loop counts as divergent: every
    // constant is then re-loaded into a vector register inside it (LDC) and turns a two-register DFMA into a
    // three-register one (3 issue cycles instead of 2, profiles/r01_fp64_issue_microbench.txt): 17 LDC and 6 more
    // three-register DFMAs per sample (round 2: 360 -> 345 instructions per sample).
#ifdef BUMP_DIVERGENT_WARP_INDEX
    const int warp = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5);
#else
    const int warp = blockIdx.x * STREAM_WARPS + __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
#endif
    if (warp >= wk.nwarps) return;
    // 32-bit group arithmetic (n_groups < 2^31 is checked on the host); 64-bit only to form addresses
    const int n_groups = (int)wk.n_groups, n_evt_groups = (int)wk.n_evt_groups, g_evt = (int)wk.g_evt;
    const int g0 = warp * (int)wk.gpw;
    const int g1 = min(g0 + (int)wk.gpw, n_groups);
    if (g0 >= g1) return;
    const int e_first = g0 < n_evt_groups ? g0 / g_evt : wk.nobs;
    double* rec = part + (size_t)rec_off[warp] * PART_STRIDE;

    // position of group g0: event e, group-in-event k
    int e = e_first;
    int k = (e < wk.nobs) ? g0 - e * g_evt : g0 - n_evt_groups;
    ThreadAcc A;
#ifdef BUMP_TMEM_ACC
    A.taddr = tmem_addr;
#endif
    acc_init(A);
    // The loads are software-pipelined at half-group granularity - y(g) is issued before x(g) is evaluated, x(g+1)
    // from inside the evaluation of y(g) - so every load has one sample evaluation (~800 cycles) to land, at no
    // extra register cost; L2 prefetches run one further group ahead.
    struct Half {
        double dl, m1, q, lm, lq, l1q, lpd;
    };
    // group g of the kernel's numbering lives in block g of the events buffer, or block g - n_evt_groups of the
    // injection buffer: the pointer advances by one block per group and is re-based once, at the set boundary
    auto block_of = [&](const int ee, const int kk) {
        return (ee >= wk.nobs ? cols.sel_base + (int64_t)kk * BLOCK_DOUBLES
                              : cols.evt_base + ((int64_t)ee * g_evt + kk) * BLOCK_DOUBLES) + lane;
    };
    auto load_half = [&](const double* p) {   // p = block + lane (x half) or block + 32 + lane (y half)
        // volatile: keeps the load where it is written (ptxas otherwise sinks it to its first use to save
        // registers, which exposes the full L2 latency once per group)
        // L2 policy evict_first: the columns are read once per evaluation, and at O5 size one pass sweeps far more than
        // the 126 MB of L2.  Marked this way they are the first lines to go, so what the other kernels of the evaluation
        // need again - the tables, the per-warp records, and the CODE of the prologue / epilogue tails - survives the
        // sweep (their single-warp tails ran twice as long after an O5-size pass: instruction fetches from DRAM).
        auto ld = [&](const double* a) {
            double v;
#ifdef BUMP_NO_L2_POLICY
            asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(a));
#else
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(l2_stream_policy));
#endif
            return v;
        };
        Half h;
        h.dl = ld(p + C_DL * GROUP);
        h.m1 = ld(p + C_M1D * GROUP);
        h.q = ld(p + C_Q * GROUP);
        h.lm = ld(p + C_LM * GROUP);
        h.lq = ld(p + C_LQ * GROUP);
        h.l1q = ld(p + C_L1Q * GROUP);
        h.lpd = ld(p + C_LPD * GROUP);
        return h;
    };
    const double* p0 = block_of(e, k);
    Half hx = load_half(p0);
    for (int g = g0; g < g1; ++g) {
        const Half hy = load_half(p0 + 32);
        auto nothing = [] {};
        if constexpr (FIXED) eval_sample_fixed<SLOT, WA>(USC_ARG hx.dl, hx.m1, hx.q, hx.lm, hx.lq, hx.l1q, hx.lpd, sb, rep, A, nothing);
        else eval_sample<SLOT, WA>(USC_ARG hx.dl, hx.m1, hx.q, hx.lm, hx.lq, hx.l1q, hx.lpd, sb, rep, A, nothing);
        // ---- advance to the next group; its x half is issued from inside the y evaluation (after y's inputs are
        // consumed), and the block after it is pulled towards L2 (28 lines: one per lane)
        int e_next = e, k_next = k + 1;
        if (k_next == ((e < wk.nobs) ? g_evt : n_groups - n_evt_groups)) {
            k_next = 0;
            e_next = e + 1;
        }
        const bool more = g + 1 < g1;
        auto next_loads = [&] {
            if (more) {
                p0 = (e_next == wk.nobs && k_next == 0) ? cols.sel_base + lane : p0 + BLOCK_DOUBLES;
                hx = load_half(p0);
                // The block after it is pulled towards L2 by a real load per line (28 lines, one per lane, result unused)
                // that carries the same evict_first policy.  Measured at O5 size (us per evaluation / tails of prologue +
                // epilogue): `prefetch.global.L2` per line 1085 / 30 (its lines enter L2 at normal priority and the pass
                // still evicts everything else), `cp.async.bulk.prefetch.L2` with the policy 1089 / 30 (no better), no
                // prefetch 1090 / 18, this 1088 / 18: the streaming kernel itself is 1.6 % slower than with the plain
                // prefetch, the serial tails 12 us shorter - a wash on one GPU, 4 % at the 8-GPU shard size.
#if defined(BUMP_L2_PREFETCH_PER_LINE)
                if (lane < BLOCK_DOUBLES / 16)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + BLOCK_DOUBLES + 15 * lane));
#elif defined(BUMP_L2_PREFETCH_BULK)
                if (lane == 0)
                    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p0 + BLOCK_DOUBLES),
                                 "n"(BLOCK_DOUBLES * 8), "l"(l2_stream_policy)
                                 : "memory");
#elif !defined(BUMP_NO_L2_PREFETCH)
                if (lane < BLOCK_DOUBLES / 16) {
                    uint32_t sink;
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;"
                                 : "=r"(sink)
                                 : "l"(p0 + BLOCK_DOUBLES + 15 * lane), "l"(l2_stream_policy));
                }
#endif
            }
        };
        if constexpr (FIXED) eval_sample_fixed<SLOT, WA>(USC_ARG hy.dl, hy.m1, hy.q, hy.lm, hy.lq, hy.l1q, hy.lpd, sb, rep, A, next_loads);
        else eval_sample<SLOT, WA>(USC_ARG hy.dl, hy.m1, hy.q, hy.lm, hy.lq, hy.l1q, hy.lpd, sb, rep, A, next_loads);
        if (!more || e_next != e) {   // event complete (for this warp): one record
            warp_flush(A, rec + (size_t)(e - e_first) * PART_STRIDE);
            acc_init(A);
        }
        e = e_next;
        k = k_next;
    }
    if (tl != nullptr && lane == 0) {   // every warp: they finish apart
        const unsigned long long t = global_ns();
        atomicMax(tl + 2 * TL_STREAM + 1, t);
        atomicMin(tl + 2 * TL_STREAM_WARPS, t);
        atomicMax(tl + 2 * TL_STREAM_WARPS + 1, t);
        if (warp < TL_WARP_SLOTS) tl[2 * TL_N + warp] = t;   // bump_debug_warp_times
    }
    // the epilogue's blocks may be scheduled once the first warp of every block is done (they then wait for the rest)
    pdl_launch_dependents<PDL_EPILOGUE && BUMP_PDL_EPI_TRIGGER == 1>();
    };   // warp_work
    warp_work();
#ifdef BUMP_TMEM_ACC
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_base), "n"(TMEM_ALLOC_COLS));
    }
#endif
}

#undef K_SC
#undef USC_PARAM
#undef USC_ARG

}  // namespace bump

// The streaming kernel: one pass over the SoA sample / injection columns producing, per tile, the
// max-shifted sums  S = sum e^{w-m},  S2 = sum e^{2(w-m)}  and the 16 gradient features  sum e^{w-m} f_k.
//
// Replaces intensity_models.py:378-381 (events) and :385-388 (injections) — z_of_dL, detector->source masses,
// LogDNDMDQDV.__call__ (:202-210), LogDNDM.__call__ (:140-151), log_smooth_turnon (:45-54), LogDNDV.__call__
// (:170-173), the Jacobian terms — plus the inner part of the logsumexp reductions (:382,389,392,401) and the
// reverse pass of all of it, in forward mode.
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
constexpr int STREAM_SMEM_BYTES = BLOB_BYTES + 16 /*mbarrier*/ + STREAM_WARPS * (NACC + 2) * 8;
constexpr double RESCALE_GAP = 60.0;   // rescale the running shift when a weight exceeds it by e^60

// ---- TMA bulk copy (global -> shared) of the table blob, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_tables(double* s_blob, uint64_t* mbar, const double* __restrict__ g_blob) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                     "r"((uint32_t)BLOB_BYTES)
                     : "memory");
        constexpr int CHUNK = 32768;
        for (int off = 0; off < BLOB_BYTES; off += CHUNK) {
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

struct SampleOut {
    double w;          // log weight without the theta-only constant; -inf if the sample carries no weight
    double f[NFEAT];
};

// One mass-function evaluation A0(m) = logaddexp(P(m), Q(m)) and its feature contributions.
__device__ __forceinline__ double mass_term(const double m, const double lm, const double* __restrict__ sc,
                                            const double2* __restrict__ mass, double (&f)[NFEAT], double& mdA) {
    const double M = sc[S_M], c = sc[S_C];
    // power-law tail with smooth turn-on, :147 and :45-54
    const double y = (m - M) * sc[S_INV_DM];
    const double e = fexp(-y);
    const double s1 = frcp(1.0 + e);
    const double sg = e * s1;                       // e/(1+e)
    const double turn = LN2 + flog(s1);             // log 2 - log1p(e)
    const double lrel = lm - sc[S_LOG_M];
    const double Q = fma(-c, lrel, sc[S_LPN] + turn);
    // PISN table lookup (:110-111) with the -inf guards of :144-145 (m <= 3 cannot happen once m >= 5)
    const bool inP = m < sc[S_TOP];
    const double pos = (m - MIN_BH_MASS) * sc[S_INV_DMBH];
    int b = __double2int_rd(pos);
    b = min(max(b, 0), NM - 2);
    const double u = pos - (double)b;
    const double2 g = mass[MR_G * NM + b];
    const double P = fma(u, g.y, g.x);
    // logaddexp(P, Q)
    const double d = inP ? (P - Q) : -INFINITY;
    const double E = fexp(-fabs(d));
    const double s = frcp(1.0 + E);
    const double Es = E * s;
    const double sP = (d > 0.0) ? s : Es;
    const double sQ = (d > 0.0) ? Es : s;
    const double A0 = ((d > 0.0) ? P : Q) - flog(s);
    // features
    const double slope = g.y * sc[S_INV_DMBH];       // dP/dm
    const double dT = sg * sc[S_INV_DM];             // d turn / dm
    mdA = sP * slope * m + sQ * fma(dT, m, -c);      // m * dA0/dm
    f[F_SQ] += sQ;
    f[F_C] = fma(sQ, lrel, f[F_C]);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double2 t = mass[(MR_GA + k) * NM + b];
        f[F_PA + k] = fma(sP, fma(u, t.y, t.x), f[F_PA + k]);
    }
    f[F_GEO] = fma(sP * slope, m - MIN_BH_MASS, f[F_GEO]);
    f[F_T] = fma(sQ * sg, m, f[F_T]);
    return A0;
}

template <bool WA>
__device__ __forceinline__ void eval_sample(const double x, const double m1d, const double q, const double lm,
                                            const double lq, const double l1q, const double lpd,
                                            const double* __restrict__ s_blob, SampleOut& o) {
    const double* __restrict__ sc = s_blob + OFF_SCAL;
    const double2* __restrict__ cos = reinterpret_cast<const double2*>(s_blob + OFF_COS);
    const double* __restrict__ ctan = s_blob + OFF_CTAN;
    const unsigned short* __restrict__ srch = reinterpret_cast<const unsigned short*>(s_blob + OFF_SRCH);
    const double2* __restrict__ mass = reinterpret_cast<const double2*>(s_blob + OFF_MASS);

    // ---- z_of_dL: b = clip(searchsorted(dl, x, side='right'), 1, n-1) - 1   (:272-273, jnp.interp)
    // bucket table keyed by the top bits of x gives a lower bound of the bin; walk up to the exact one.
    int j = (__double2hiint(x) >> (20 - SRCH_MBITS)) - (SRCH_EXP_LO << SRCH_MBITS);
    j = min(max(j, 0), SRCH_N - 1);
    int b = srch[j];
    while (b < NZ - 2 && x >= cos[CR_DL * NZ + b + 1].x) ++b;
    const bool beyond = x > sc[S_DL_LAST];          // jnp.interp clamps to fp[-1]; no gradient flows to x or xp
    const double2 rdl = cos[CR_DL * NZ + b];
    double t = (x - rdl.x) * rdl.y;
    const double idl = beyond ? 0.0 : rdl.y;
    t = beyond ? 1.0 : t;
    // ---- position inside the z bin: 1+z = (1+z_b)(1 + t eps)
    const double2 rz = cos[CR_Z * NZ + b];
    const double zeps = sc[S_ZEPS];
    const double te = t * zeps;
    const double u1 = frcp1p_small(te);
    const double ropz = rz.x * u1;                  // 1/(1+z)
    const double L = rz.y + flog1p_small(te);       // log1p(z)
    const double lt = zeps * u1;                    // dL/dt
    // ---- source-frame masses (:379, :207)
    double m1 = m1d * ropz;
    double m2 = q * m1;
    double lm1 = lm - L;
    double lm2 = lm1 + lq;
    bool valid = (m1 >= MBH_MIN) && (m2 >= MBH_MIN);   // :149
    if (!valid) {   // keep every intermediate finite; the sample gets zero weight below
        m1 = MREF; m2 = MREF; lm1 = LOG_MREF; lm2 = LOG_MREF;
    }
    // ---- dVC/dz and d(dL)/dz lerps at z (:264-268), same bin, same t
    const double2 rvc = cos[CR_DVC * NZ + b];
    const double2 rdd = cos[CR_DDL * NZ + b];
    const double dvc = fma(t, rvc.y, rvc.x);
    const double ddl = fma(t, rdd.y, rdd.x);
    valid = valid && (dvc > 0.0);
    const double idvc = frcp(valid ? dvc : 1.0);
    const double iddl = frcp(ddl);
    const double ljac = flog((valid ? dvc : 1.0) * iddl);
    // ---- merger-rate density (:173)
    const double lam = sc[S_LAM], kappa = sc[S_KAPPA], beta = sc[S_BETA];
    const double r = fexp(kappa * (L - sc[S_LOPZP]));
    const double sr = frcp(1.0 + r);
    const double sig = r * sr;
    const double V0 = fma(lam, L, flog(sr));        // lam*log1p(z) - log1p(r)
#pragma unroll
    for (int k = 0; k < NFEAT; ++k) o.f[k] = 0.0;
    double mdA1, mdA2;
    const double A1 = mass_term(m1, lm1, sc, mass, o.f, mdA1);
    const double A2 = mass_term(m2, lm2, sc, mass, o.f, mdA2);
    const double pair = lm1 + l1q;                  // log(m1+m2); the -log(60) is in the constant
    double w = A1 + A2 + fma(beta, pair, lm1) + V0 - 2.0 * L + ljac - lpd;   // :210, :381
    // ---- d w / d t at fixed tables, then the cosmological tangents
    const double Wt = lt * (lam - kappa * sig - 3.0 - beta - mdA1 - mdA2) + rvc.y * idvc - rdd.y * iddl;
    const double Wx = Wt * idl;                     // -Wx * (d dl-table / d theta)(t) = (dw/dt)(dt/dtheta)
    o.f[F_CZ] = Wx * x;
    auto tangent = [&](const int tdl, const int tdvc, const int tddl) -> double {
        const double a0 = ctan[tdl * NZ + b], a1 = ctan[tdl * NZ + b + 1];
        const double v0 = ctan[tdvc * NZ + b], v1 = ctan[tdvc * NZ + b + 1];
        const double d0 = ctan[tddl * NZ + b], d1 = ctan[tddl * NZ + b + 1];
        return fma(-Wx, fma(t, a1 - a0, a0), fma(t, v1 - v0, v0) * idvc - fma(t, d1 - d0, d0) * iddl);
    };
    o.f[F_OM] = tangent(CT_DL_OM, CT_DVC_OM, CT_DDL_OM);
    o.f[F_W] = tangent(CT_DL_W, CT_DVC_W, CT_DDL_W);
    if constexpr (WA) o.f[F_WA] = tangent(CT_DL_WA, CT_DVC_WA, CT_DDL_WA);
    o.f[F_BETA] = pair;
    o.f[F_L] = L;
    o.f[F_SIG] = sig;
    o.f[F_SIGL] = sig * L;
    o.w = valid ? w : -INFINITY;
}

struct ThreadAcc {
    double m;           // running shift (max-like)
    double a[NACC];     // S, S2, features
    double nvalid;
};

__device__ __forceinline__ void acc_init(ThreadAcc& A) {
    A.m = -INFINITY;
    A.nvalid = 0.0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) A.a[k] = 0.0;
}

__device__ __forceinline__ void acc_add(ThreadAcc& A, const SampleOut& o) {
    if (o.w == -INFINITY) return;
    if (o.w - A.m > RESCALE_GAP) {   // also the first finite sample (m = -inf)
        const double sc = (A.m == -INFINITY) ? 0.0 : fexp(A.m - o.w);
        A.a[0] *= sc;
        A.a[1] *= sc * sc;
#pragma unroll
        for (int k = 2; k < NACC; ++k) A.a[k] *= sc;
        A.m = o.w;
    }
    const double p = fexp(o.w - A.m);
    A.a[0] += p;
    A.a[1] = fma(p, p, A.a[1]);
#pragma unroll
    for (int k = 0; k < NFEAT; ++k) A.a[2 + k] = fma(p, o.f[k], A.a[2 + k]);
    A.nvalid += 1.0;
}

// Block-wide merge of the per-thread accumulators, deterministic order; result written by thread 0..NACC+1.
__device__ __forceinline__ void tile_reduce(ThreadAcc& A, double* red, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double mx = A.m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();   // protects `red` against the previous tile's readers
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < STREAM_WARPS; ++w) mx = fmax(mx, red[w]);
    const double sc = (A.m == -INFINITY) ? 0.0 : fexp(A.m - mx);
    A.a[0] *= sc;
    A.a[1] *= sc * sc;
#pragma unroll
    for (int k = 2; k < NACC; ++k) A.a[k] *= sc;
    double v[NACC + 1];
#pragma unroll
    for (int k = 0; k < NACC; ++k) v[k] = A.a[k];
    v[NACC] = A.nvalid;
#pragma unroll
    for (int k = 0; k <= NACC; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();
    double* r2 = red + STREAM_WARPS;   // [STREAM_WARPS][NACC+1]
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k <= NACC; ++k) r2[warp * (NACC + 1) + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x <= NACC) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < STREAM_WARPS; ++w) s += r2[w * (NACC + 1) + threadIdx.x];
        out[1 + threadIdx.x] = s;   // [1..NACC] sums, [NACC+1] nvalid
    }
    if (threadIdx.x == 0) out[0] = mx;
}

template <bool WA>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
stream_kernel(const Columns cols, const Tile* __restrict__ tiles, const int ntiles,
              const double* __restrict__ g_blob, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* s_blob = reinterpret_cast<double*>(smem_raw);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + BLOB_BYTES);
    double* red = reinterpret_cast<double*>(smem_raw + BLOB_BYTES + 16);

    stage_tables(s_blob, mbar, g_blob);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const Tile T = tiles[tile];
        const double* const* c = T.set ? cols.sel : cols.evt;
        const double2* c_dl = reinterpret_cast<const double2*>(c[C_DL] + T.off);
        const double2* c_m1 = reinterpret_cast<const double2*>(c[C_M1D] + T.off);
        const double2* c_q = reinterpret_cast<const double2*>(c[C_Q] + T.off);
        const double2* c_lm = reinterpret_cast<const double2*>(c[C_LM] + T.off);
        const double2* c_lq = reinterpret_cast<const double2*>(c[C_LQ] + T.off);
        const double2* c_l1q = reinterpret_cast<const double2*>(c[C_L1Q] + T.off);
        const double2* c_lpd = reinterpret_cast<const double2*>(c[C_LPD] + T.off);
        ThreadAcc A;
        acc_init(A);
        const int npairs = T.count >> 1;
        for (int p = threadIdx.x; p < npairs; p += STREAM_THREADS) {
            const double2 dl = __ldg(c_dl + p), m1 = __ldg(c_m1 + p), q = __ldg(c_q + p), lm = __ldg(c_lm + p),
                          lq = __ldg(c_lq + p), l1q = __ldg(c_l1q + p), lpd = __ldg(c_lpd + p);
            SampleOut o;
            eval_sample<WA>(dl.x, m1.x, q.x, lm.x, lq.x, l1q.x, lpd.x, s_blob, o);
            acc_add(A, o);
            eval_sample<WA>(dl.y, m1.y, q.y, lm.y, lq.y, l1q.y, lpd.y, s_blob, o);
            acc_add(A, o);
        }
        tile_reduce(A, red, part + (size_t)tile * PART_STRIDE);
    }
}

}  // namespace bump

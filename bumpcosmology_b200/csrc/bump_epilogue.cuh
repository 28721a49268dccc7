// Epilogue: merge the per-tile partials of the streaming kernel into this rank's partial, and finalize the
// (rank-merged) partials into the result header.
//
//   epilogue_kernel : per event, merge its tiles -> logsumexp (intensity_models.py:382), Neff (:401), normalised
//                     gradient features; sum over events in a fixed order; merge the injection tiles (:389,392).
//   finalize        : rank-ordered merge of PARTIAL_LEN-double partials, then the theta-only constants, the
//                     selection statistics (:389-394) and the chain rule from features to d/dtheta.
//                     `finalize_merge` is __host__ __device__: the same code backs bump_merge_partials (host) and
//                     finalize_kernel (device), so every rank computes identical bits.
#pragma once
#include <math.h>

#include "bump_layout.cuh"

namespace bump {

constexpr int EPI_THREADS = 256;

// Gradient of (logsumexp + theta-only constant) from softmax-weighted features phi (already normalised and, for
// events, summed over n events).  g[15] in theta order; sc = scalar block of the table blob.
// (not inlined on the device: it runs twice in the single-thread tail of the epilogue, whose instructions come from a
// cold instruction cache after a streaming kernel that swept L2 - code size is latency there)
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
void grad_from_features(const double* phi, const double n, const double* sc, double* g) {
    const double inv_h = sc[S_INV_H];
    const double c = sc[S_C], M = sc[S_M], kappa = sc[S_KAPPA], zp = sc[S_ZP];
    const double sq = phi[F_SQ];
    const double geo = phi[F_GEO] * sc[S_INV_TOPM3];
    const double* lpn = sc + S_LPN_D0;   // d log_pl_norm / d(a, b, mpisn, mbhmax, sigma)
    const double* ln = sc + S_LN_D0;     // d log_norm / d(a, b, c, mpisn, mbhmax, sigma, fpl)
    // cosmology: d_L, dd_L ~ 1/h and dV_C ~ 1/h^3 (intensity_models.py:231-239)
    g[T_H] = (phi[F_CZ] - 2.0 * n) * inv_h;
    g[T_OM] = phi[F_OM];
    g[T_W] = phi[F_W];
    // mass function (two evaluations per sample, each normalised by log_norm)
    g[T_A] = phi[F_PA] + lpn[0] * sq + 2.0 * n * ln[0];
    g[T_B] = phi[F_PB] + lpn[1] * sq + 2.0 * n * ln[1];
    g[T_C] = -phi[F_C] + 2.0 * n * ln[2];
    g[T_MPISN] = phi[F_PMPISN] + lpn[2] * sq + 2.0 * n * ln[3];
    g[T_MBHMAX] = phi[F_PMBHMAX] - geo + sq * (c / M + lpn[3]) - phi[F_T] * sc[S_INV_DM] / M + 2.0 * n * ln[4];
    g[T_SIGMA] = phi[F_PSIGMA] - 7.0 * geo + lpn[4] * sq + 2.0 * n * ln[5];
    g[T_FPL] = sq / sc[S_FPL] + 2.0 * n * ln[6];
    g[T_BETA] = phi[F_BETA] - n * LOG_MREF_PAIR;
    // rate (intensity_models.py:167-173)
    g[T_LAM] = phi[F_L];
    g[T_KAPPA] = -(phi[F_SIGL] - sc[S_LOPZP] * phi[F_SIG]) + n * sc[S_LNV_KAPPA];
    g[T_ZP] = phi[F_SIG] * kappa / (1.0 + zp) + n * sc[S_LNV_ZP];
    g[T_WA] = (sc[S_USE_WA] != 0.0) ? phi[F_WA] : 0.0;
    if (sc[S_FIXED] != 0.0) g[T_H] = g[T_OM] = g[T_W] = g[T_WA] = 0.0;   // pop_model: the cosmology is not a parameter
}

// partials: [nranks][PARTIAL_LEN] in rank order -> out[OUT_HEADER], in two independent halves (the last block of the
// epilogue runs them in two warps at once; the host-side merge calls them one after the other): the events' factor ...
__host__ __device__ inline void finalize_events(const double* partials, const int nranks, double* out) {
    const double* sc = partials + P_SCAL0;   // identical on every rank (same theta)
    double llsum = 0.0, nobs = 0.0, nvalid_e = 0.0, ndead = 0.0;
    double phi[NFEAT];
    for (int k = 0; k < NFEAT; ++k) phi[k] = 0.0;
    for (int r = 0; r < nranks; ++r) {
        const double* p = partials + (size_t)r * PARTIAL_LEN;
        llsum += p[P_LLSUM];
        nobs += p[P_NOBS];
        nvalid_e += p[P_NVALID_EVT];
        ndead += p[P_NDEAD_EVT];
        for (int k = 0; k < NFEAT; ++k) phi[k] += p[P_FSUM0 + k];
    }
    // 'loglike' factor (:382-383)
    out[OUT_LOGLIKE] = (ndead > 0.0) ? -INFINITY : llsum + nobs * (sc[S_CONST] - sc[S_LOG_NSAMP]);
    grad_from_features(phi, nobs, sc, out + OUT_DLOGLIKE);
    out[OUT_NVALID_EVT] = nvalid_e;
    out[OUT_NOBS] = nobs;
}

// ... and the injections' (:389-394)
__host__ __device__ inline void finalize_selection(const double* partials, const int nranks, double* out) {
    const double* sc = partials + P_SCAL0;
    double nvalid_s = 0.0, nsel = 0.0;
    double M = -INFINITY;
    for (int r = 0; r < nranks; ++r) {
        const double* p = partials + (size_t)r * PARTIAL_LEN;
        nvalid_s += p[P_NVALID_SEL];
        nsel += p[P_NSEL];
        M = fmax(M, p[P_SEL_M]);
    }
    double acc[NACC];
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int r = 0; r < nranks; ++r) {
        const double* p = partials + (size_t)r * PARTIAL_LEN;
        if (p[P_SEL_M] == -INFINITY) continue;
        const double s = exp(p[P_SEL_M] - M);
        acc[0] += p[P_SEL_ACC0] * s;
        acc[1] += p[P_SEL_ACC0 + 1] * (s * s);
        for (int k = 2; k < NACC; ++k) acc[k] += p[P_SEL_ACC0 + k] * s;
    }
    const double cst = sc[S_CONST];
    const double lnd = sc[S_LOG_NDRAW];
    const double log_mu = M + log(acc[0]) + cst - lnd;
    const double log_mu2 = 2.0 * M + log(acc[1]) + 2.0 * cst - 2.0 * lnd;
    const double log_s2 = log_mu2 + log1p(-exp(2.0 * log_mu - lnd - log_mu2));
    out[OUT_LOG_MU_SEL] = log_mu;
    out[OUT_LOG_MU2] = log_mu2;
    out[OUT_NEFF_SEL] = exp(2.0 * log_mu - log_s2);
    double phis[NFEAT];
    const double iS = 1.0 / acc[0];
    for (int k = 0; k < NFEAT; ++k) phis[k] = acc[2 + k] * iS;
    grad_from_features(phis, 1.0, sc, out + OUT_DLOG_MU);
    out[OUT_NVALID_SEL] = nvalid_s;
    out[OUT_NSEL] = nsel;
}

// theta outside the support: the reference yields NaN (NUTS treats it as divergent)
__host__ __device__ inline bool finalize_is_bad(const double* partials) { return partials[P_SCAL0 + S_BAD] != 0.0; }

__host__ __device__ inline void finalize_merge(const double* partials, const int nranks, double* out) {
    for (int k = 0; k < OUT_HEADER; ++k) out[k] = 0.0;
    finalize_events(partials, nranks, out);
    finalize_selection(partials, nranks, out);
    if (finalize_is_bad(partials)) {
        for (int k = 0; k < OUT_NVALID_EVT; ++k) out[k] = NAN;
    }
}

// ---- fused exchange over peer memory (NVLink): every rank owns a mailbox that all peers can write.
// The last block of the epilogue pushes this rank's 1 KiB partial into every mailbox (its own included), waits until
// its own mailbox holds the current evaluation's partial of every rank, and finalizes - no NCCL call, no extra launch.
//
// The transport is a low-latency line protocol (the idea of NCCL's LL): every double travels as ONE 16-byte line
// {low word, epoch, high word, epoch}, written with a single 16-byte store.  Each 8-byte half carries its own copy of
// the 32-bit epoch, so the receiver needs nothing but 8-byte store atomicity: it polls the line until both epochs are
// the current one, and then holds the data.  No fence between data and flag, no flag round trip, no second read of the
// mailbox: one NVLink one-way latency from "partial ready" to "partial received" (the first version - data, system
// fence, flag, poll, system fence, read - cost two NVLink round trips more: ~5 us of every multi-rank evaluation).
// line[parity][rank][k]: two parities, because a rank can run at most one evaluation ahead of the slowest peer (it
// needs that peer's current partial to finish), so the lines it overwrites have been consumed; a line of parity p is
// next written two epochs later, and its stale epoch never equals the awaited one.
//
// Failure is COLLECTIVE and STICKY.  A rank that waits longer than `timeout_ns` for a peer gives up, does NOT advance
// its epoch, marks its exchange state broken and sets flag[both parities][its rank] in every peer's mailbox to
// P2P_POISON; a peer that is waiting for its lines - or arrives later, however late - finds the poison, fails the same
// way and poisons everybody else in turn.  Every evaluation after that fails immediately until the ranks detach and
// attach again.  The result header carries the status word (OUT_STATUS) next to NaN outputs, and bump_eval returns
// BUMP_E_EXCHANGE: ranks can no longer disagree about whether an evaluation happened.
constexpr int P2P_MAX_RANKS = 16;
constexpr unsigned long long P2P_POISON = ~0ull;
struct Mailbox {
    uint4 line[2][P2P_MAX_RANKS][PARTIAL_LEN];        // {lo, epoch32, hi, epoch32}
    unsigned long long flag[2][P2P_MAX_RANKS];        // P2P_POISON: that rank's exchange has failed
};
struct Peers {
    Mailbox* box[P2P_MAX_RANKS];   // box[r] = rank r's mailbox as mapped into this process (box[rank] = local)
    int nranks, rank;
    unsigned long long timeout_ns;
};
// exchange state of a context: [0] epoch of the last completed exchange, [1] broken (sticky)

// Called by all threads of ONE block.  `parts` (shared memory, [nranks][PARTIAL_LEN]) holds this rank's partial in
// its first PARTIAL_LEN entries on entry and every rank's partial in rank order on a successful return.  Returns
// STATUS_OK, STATUS_EXCHANGE_TIMEOUT (a peer never arrived) or STATUS_EXCHANGE_POISONED (a peer failed, now or earlier).
__device__ inline double p2p_exchange(const Peers& peers, double* parts, unsigned long long* __restrict__ state) {
    __shared__ int s_status;
    const int tid = threadIdx.x;
    const unsigned long long epoch = state[0] + 1ull;
    const bool broken = state[1] != 0ull;
    const int par = (int)(epoch & 1ull);
    const uint32_t e32 = (uint32_t)epoch;
    const int n = peers.nranks * PARTIAL_LEN;
    if (tid == 0) s_status = broken ? 2 : 0;
    // A peer that failed EARLIER (it timed out waiting for this rank, say) has left its poison here, but its lines of
    // this epoch may be complete all the same - it sends before it waits.  So the poison flags are looked at whether or
    // not the data arrives; the load is issued now and consumed after the sends.
    unsigned long long poisoned = 0ull;
    if (!broken && tid < peers.nranks)
        poisoned = *reinterpret_cast<const volatile unsigned long long*>(&peers.box[peers.rank]->flag[par][tid]);
    if (!broken) {
        for (int idx = tid; idx < n; idx += blockDim.x) {
            const int r = idx / PARTIAL_LEN, k = idx % PARTIAL_LEN;
            const double v = parts[k];
            uint4* dst = &peers.box[r]->line[par][peers.rank][k];
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((uint32_t)__double2loint(v)),
                         "r"(e32), "r"((uint32_t)__double2hiint(v)), "r"(e32)
                         : "memory");
        }
    }
    __syncthreads();   // every thread has read this rank's partial out of `parts`; s_status is set
    if (poisoned == P2P_POISON) atomicMax(&s_status, 2);
    if (!broken) {
        const unsigned long long t0 = global_ns();
        for (int idx = tid; idx < n; idx += blockDim.x) {
            const int r = idx / PARTIAL_LEN, k = idx % PARTIAL_LEN;
            const uint4* src = &peers.box[peers.rank]->line[par][r][k];
            const volatile unsigned long long* poison = &peers.box[peers.rank]->flag[par][r];
            uint32_t lo = 0u, e0 = 0u, hi = 0u, e1 = 0u;
            for (unsigned int spins = 1;; ++spins) {
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lo), "=r"(e0), "=r"(hi), "=r"(e1)
                             : "l"(src)
                             : "memory");
                if (e0 == e32 && e1 == e32) break;
                if ((spins & 7u) == 0u) {   // the failure checks stay off the fast path
                    if (*reinterpret_cast<volatile int*>(&s_status) != 0) break;   // another thread has given up
                    if (*poison == P2P_POISON) {
                        atomicMax(&s_status, 2);
                        break;
                    }
                    if (global_ns() - t0 > peers.timeout_ns) {   // a peer is gone: fail instead of hanging the GPU
                        atomicMax(&s_status, 1);
                        break;
                    }
                }
            }
            parts[idx] = __hiloint2double((int)hi, (int)lo);
        }
    }
    __syncthreads();
    const int st = s_status;
    if (st != 0) {
        if (tid < peers.nranks && tid != peers.rank) {   // tell every peer, whichever evaluation it is in
            *reinterpret_cast<volatile unsigned long long*>(&peers.box[tid]->flag[0][peers.rank]) = P2P_POISON;
            *reinterpret_cast<volatile unsigned long long*>(&peers.box[tid]->flag[1][peers.rank]) = P2P_POISON;
        }
        if (tid == 0) state[1] = 1ull;
        __threadfence_system();
    } else if (tid == 0) {
        state[0] = epoch;
    }
    __syncthreads();
    return st == 0 ? STATUS_OK : st == 1 ? STATUS_EXCHANGE_TIMEOUT : STATUS_EXCHANGE_POISONED;
}

__global__ void finalize_kernel(const double* __restrict__ partials, const int nranks, double* __restrict__ out,
                                unsigned long long* __restrict__ tl) {
    timeline_begin(tl, TL_FINALIZE);
    if (threadIdx.x == 0 && blockIdx.x == 0) finalize_merge(partials, nranks, out);
    timeline_end(tl, TL_FINALIZE);
}

// Merge two max-shifted accumulators (m, a[NACC]) <- (m, a) (+) (m2, b[NACC]).
__device__ __forceinline__ void lse_merge(double& m, double* a, const double m2, const double* b) {
    if (m2 == -INFINITY) return;
    const double mx = fmax(m, m2);
    const double s1 = (m == -INFINITY) ? 0.0 : exp(m - mx);
    const double s2 = exp(m2 - mx);
    a[0] = a[0] * s1 + b[0] * s2;
    a[1] = a[1] * (s1 * s1) + b[1] * (s2 * s2);
#pragma unroll
    for (int k = 2; k < NACC; ++k) a[k] = a[k] * s1 + b[k] * s2;
    m = mx;
}

// Deterministic block sum of NV values per thread (fixed shuffle tree, then fixed-order sum over warps).
template <int NV>
__device__ __forceinline__ void epi_block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
        for (int w = 0; w < EPI_THREADS / 32; ++w) s += red[w * NV + k];
        v[k] = s;
    }
}

// part: [nrecords][PART_STRIDE] written by the streaming kernel; event e (groups [e*g_evt, (e+1)*g_evt)) is covered
// by warps floor(e*g_evt/gpw) .. floor(((e+1)*g_evt-1)/gpw), merged here in that fixed order.
//   blocks 0 .. nb_evt-1 : one thread per event -> logsumexp, Neff, normalised features; block sum -> slot[b]
//   blocks nb_evt .. nb_evt+nb_sel-1 : a slice of the injection records each -> (shift, sums) in slot[b]
//   last block to finish : fixed-order sum of the slots -> this rank's partial; with `out_header` non-null
//                          (single rank) it also finalizes, saving a launch.
constexpr int EPI_SLOT = 24;
__global__ void __launch_bounds__(EPI_THREADS)
epilogue_kernel(const double* __restrict__ part, const int* __restrict__ rec_off, const Work wk, const double nsel,
                const int lpe /* lanes per event: power of two <= 32 */, const int nb_sel /* injection blocks */,
                const double* __restrict__ blob,
                double* __restrict__ neff_out, double* __restrict__ slots, unsigned int* __restrict__ ticket,
                double* __restrict__ partial, double* __restrict__ out_header, const Peers* __restrict__ peers,
                unsigned long long* __restrict__ exchange_state, unsigned long long* __restrict__ tl,
                unsigned long long* __restrict__ done_seq /* evaluations completed (host call's graph), or null */) {
    __shared__ double red[32 + (EPI_THREADS / 32) * 32];
    static_assert(32 + (EPI_THREADS / 32) * 32 >= (EPI_THREADS / 32) * (NACC + 3), "block sums fit");
    __shared__ double s_max;
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    pdl_wait<PDL_EPILOGUE>();   // scheduled while the stream kernel drains: its records (and the prologue's scalars) are complete after this
    timeline_begin(tl, TL_EPILOGUE);
    const int nobs = wk.nobs;
    const int epb = EPI_THREADS / lpe;   // events per block
    const int nb_evt = (nobs + epb - 1) / epb;
    double* slot = slots + (size_t)blockIdx.x * EPI_SLOT;
    // 32-bit group arithmetic (n_groups < 2^31 is checked on the host; 64-bit divisions are emulated and slow)
    const int g_evt = (int)wk.g_evt, gpw = (int)wk.gpw, n_evt_groups = (int)wk.n_evt_groups, n_groups = (int)wk.n_groups;
    auto rec_index = [&](const int w, const int e) {   // warp w's records start at rec_off[w], ordered by event
        const int g0 = w * gpw;
        return rec_off[w] + (e - (g0 < n_evt_groups ? g0 / g_evt : nobs));
    };
    if ((int)blockIdx.x < nb_evt) {
        // ---- events
        double ev[NFEAT + 3];   // llsum, nvalid, ndead, phi[NFEAT]
#pragma unroll
        for (int k = 0; k < NFEAT + 3; ++k) ev[k] = 0.0;
        const int e = blockIdx.x * epb + tid / lpe;
        const int sub = tid % lpe;
        double m = -INFINITY, a[NACC], nv = 0.0;
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.0;
        if (e < nobs) {
            const int w0 = (e * g_evt) / gpw, w1 = ((e + 1) * g_evt - 1) / gpw;
            for (int w = w0 + sub; w <= w1; w += lpe) {   // one record per lane unless the event spans > 32 warps
                const double* p = part + (size_t)rec_index(w, e) * PART_STRIDE;
                double b[NACC];
#pragma unroll
                for (int k = 0; k < NACC; ++k) b[k] = p[1 + k];
                lse_merge(m, a, p[0], b);
                nv += p[1 + NACC];
            }
        }
        // The event's lanes: common shift first (max over the lanes), every lane rescales its own sums once, then a
        // plain xor butterfly of additions - a + b is commutative, so both partners of every step hold identical bits
        // and the result is deterministic.  (Round 1 merged (shift, sums) pairs at every step: two exponentials and a
        // select chain per step instead of one exponential in total.)
        double mx = m;
        for (int o = lpe >> 1; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        {
            const double sc = (m == -INFINITY) ? 0.0 : exp(m - mx);
            a[0] *= sc;
            a[1] *= sc * sc;
#pragma unroll
            for (int k = 2; k < NACC; ++k) a[k] *= sc;
        }
        for (int o = lpe >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
            nv += __shfl_xor_sync(0xffffffffu, nv, o);
        }
        m = mx;
        if (e < nobs && sub == 0) {
            if (m == -INFINITY || !(a[0] > 0.0)) {   // no finite-weight sample: logsumexp = -inf (reference: same)
                ev[2] = 1.0;
                neff_out[e] = NAN;
            } else {
                const double iS = 1.0 / a[0];
                ev[0] = m + log(a[0]);
                ev[1] = nv;
                neff_out[e] = a[0] * a[0] / a[1];   // exp(2 lse(w) - lse(2w)), :401
#pragma unroll
                for (int k = 0; k < NFEAT; ++k) ev[3 + k] = a[2 + k] * iS;
            }
        }
        epi_block_sum<NFEAT + 3>(ev, red);
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < NFEAT + 3; ++k) slot[k] = ev[k];
        }
    } else {
        // ---- injections (pseudo-event nobs): global shift first, then plain sums over the warps owning its groups
        // (the warps that own injection groups are split evenly over the nb_sel injection blocks: one block was the
        // longest-running of the kernel at GWTC-3 size)
        const bool has_sel = n_groups > n_evt_groups;
        const int wa0 = has_sel ? n_evt_groups / gpw : 0;
        const int wa1 = has_sel ? (n_groups - 1) / gpw : -1;
        const int per = (wa1 - wa0 + nb_sel) / nb_sel;                 // ceil((wa1 - wa0 + 1) / nb_sel)
        const int ws0 = wa0 + ((int)blockIdx.x - nb_evt) * per;
        const int ws1 = min(wa1, ws0 + per - 1);
        double mx = -INFINITY;
        for (int w = ws0 + tid; w <= ws1; w += EPI_THREADS)
            mx = fmax(mx, part[(size_t)rec_index(w, nobs) * PART_STRIDE]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            double m = red[0];
            for (int w = 1; w < EPI_THREADS / 32; ++w) m = fmax(m, red[w]);
            s_max = m;
        }
        __syncthreads();
        mx = s_max;
        double sv[NACC + 1];
#pragma unroll
        for (int k = 0; k <= NACC; ++k) sv[k] = 0.0;
        for (int w = ws0 + tid; w <= ws1; w += EPI_THREADS) {
            const double* p = part + (size_t)rec_index(w, nobs) * PART_STRIDE;
            if (p[0] == -INFINITY) continue;
            const double s = exp(p[0] - mx);
            sv[0] += p[1] * s;
            sv[1] += p[2] * (s * s);
#pragma unroll
            for (int k = 2; k < NACC; ++k) sv[k] += p[1 + k] * s;
            sv[NACC] += p[1 + NACC];
        }
        epi_block_sum<NACC + 1>(sv, red);
        if (tid == 0) {
            slot[0] = mx;
#pragma unroll
            for (int k = 0; k <= NACC; ++k) slot[1 + k] = sv[k];
        }
    }
    // ---- last block: fixed-order sum of the slots
    __threadfence();
    __syncthreads();
    timeline_begin(tl, TL_EPI_BLOCKS);   // (its minimum is meaningless: only the end of the phase is of interest)
    timeline_end(tl, TL_EPI_BLOCKS);
    if (tid == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) {
        timeline_end(tl, TL_EPILOGUE);
        return;
    }
    __threadfence();
    timeline_begin(tl, TL_EPI_LAST);
    __shared__ double s_part[P2P_MAX_RANKS * PARTIAL_LEN];
    __shared__ double s_out[OUT_HEADER];
    __shared__ double s_sel[8 * EPI_SLOT];   // the injection blocks' (shift, sums)
    // Everything this block reads from L2 is requested up front, so that the round trips overlap: the injection
    // blocks' (shift, sums), the scalar block (both consumed further down) and the event blocks' slots.
    static_assert(8 * EPI_SLOT <= EPI_THREADS, "one thread per staged value");
    const double sel_v = (tid < nb_sel * EPI_SLOT) ? __ldcg(slots + (size_t)nb_evt * EPI_SLOT + tid) : 0.0;
    const double scal_v = (tid >= P_SCAL0 && tid < P_SCAL0 + NSCAL) ? __ldcg(blob + OFF_SCAL + tid - P_SCAL0) : 0.0;
    {   // column k of the slots, summed by 8 threads (blocks p, p + 8, ...) and then over p in a fixed order
        const int k = tid & 31, p = tid >> 5;
        static_assert(NFEAT + 3 <= 32 && EPI_THREADS == 256, "8 x 32 threads cover the slot columns");
        double s = 0.0;
        if (k < NFEAT + 3) {
#pragma unroll 4
            for (int b = p; b < nb_evt; b += EPI_THREADS / 32) s += __ldcg(slots + (size_t)b * EPI_SLOT + k);
        }
        __syncthreads();   // `red` was last used by the block sums above
        red[32 + p * 32 + k] = s;
        if (tid < nb_sel * EPI_SLOT) s_sel[tid] = sel_v;
        __syncthreads();
        if (tid < NFEAT + 3) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < EPI_THREADS / 32; ++q) t += red[32 + q * 32 + tid];
            red[tid] = t;
        }
    }
    __syncthreads();
    // this rank's partial: to global memory (host-driven / NCCL exchanges read it there) and to shared memory, from
    // which one thread finalizes - its ~250 dependent operations then run on 30-cycle shared-memory loads instead of
    // L2 round trips (the serial tail was half of the epilogue's time at GWTC-3 size)
    if (tid < P_SCAL0 + NSCAL) {
        double v = 0.0;
        if (tid < P_SCAL0) {
            if (tid == P_LLSUM) v = red[0];
            else if (tid == P_NOBS) v = (double)nobs;
            else if (tid >= P_FSUM0 && tid < P_FSUM0 + NFEAT) v = red[3 + tid - P_FSUM0];
            else if (tid == P_NVALID_EVT) v = red[1];
            else if (tid == P_NDEAD_EVT) v = red[2];
            else if (tid == P_SEL_M || (tid >= P_SEL_ACC0 && tid < P_SEL_ACC0 + NACC) || tid == P_NVALID_SEL) {
                // rank-order merge of the injection blocks' (shift, sums): common shift, then a fixed-order sum
                double M = -INFINITY;
                for (int b = 0; b < nb_sel; ++b) M = fmax(M, s_sel[b * EPI_SLOT]);
                if (tid == P_SEL_M) {
                    v = M;
                } else {
                    const int k = (tid == P_NVALID_SEL) ? NACC : tid - P_SEL_ACC0;
                    for (int b = 0; b < nb_sel; ++b) {
                        const double mb = s_sel[b * EPI_SLOT];
                        if (mb == -INFINITY) continue;
                        const double sc = (k == NACC) ? 1.0 : exp(mb - M);
                        v += s_sel[b * EPI_SLOT + 1 + k] * (k == 1 ? sc * sc : sc);
                    }
                }
            }
            else if (tid == P_NSEL) v = nsel;
        } else {
            v = scal_v;
        }
        partial[tid] = v;
        s_part[tid] = v;
    }
    static_assert(P_SCAL0 + NSCAL == PARTIAL_LEN && PARTIAL_LEN <= EPI_THREADS, "one thread per entry of the partial");
    if (tid == 0) *ticket = 0u;
    if (out_header) {
        int nparts = 1;
        double status = STATUS_OK;
        if (peers) {   // multi-rank, fused exchange over peer memory: s_part <- every rank's partial
            __syncthreads();
            status = p2p_exchange(*peers, s_part, exchange_state);
            if (status == STATUS_OK) nparts = peers->nranks;
        }
        if (tid < OUT_HEADER) s_out[tid] = 0.0;
        __syncthreads();
        // the two factors are independent chains of ~120 dependent operations each: one thread of warp 0 takes the
        // events', one of warp 1 the injections'
        if (status == STATUS_OK) {
            if (tid == 0) finalize_events(s_part, nparts, s_out);
            if (tid == 32) finalize_selection(s_part, nparts, s_out);
        }
        __syncthreads();
        if (tid < OUT_HEADER) {
            double v = s_out[tid];
            if (status != STATUS_OK) {   // the exchange failed on every rank (see p2p_exchange): no result, and say so
                v = (tid < OUT_NVALID_EVT) ? NAN : (tid == OUT_STATUS ? status : 0.0);
            } else if (finalize_is_bad(s_part) && tid < OUT_NVALID_EVT) {
                v = NAN;
            }
            out_header[tid] = v;
        }
        // the host call's graph copies this counter to pinned host memory behind the result: the calling thread polls
        // it there instead of synchronising the stream
        if (done_seq != nullptr && tid == 0) *done_seq = *done_seq + 1ull;
    }
    timeline_end(tl, TL_EPI_LAST);
    timeline_end(tl, TL_EPILOGUE);
}

}  // namespace bump

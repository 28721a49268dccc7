// Prologue: ONE kernel per evaluation (prologue_kernel) builds the theta-dependent tables with forward-mode tangents
// (256 PISN rows + 4 cosmology chunks, one block each, launched as thread-block clusters of 4) and then, in whichever
// block finishes last, packs the per-bin records, the d_L bucket table and the scalars that the streaming kernel
// bulk-copies into shared memory.  (Round 1 used two launches and a spin chain in global memory for the cumulative
// trapezoid; the chain now runs over distributed shared memory inside the cosmology cluster, whose four blocks the
// hardware co-schedules, so it cannot deadlock and needs no flag to re-arm.)
#pragma once
#include <cooperative_groups.h>

#include "bump_dual.cuh"
#include "bump_layout.cuh"

namespace bump {

// aux (global) workspace layout, doubles: raw knots and tangents (also what bump_debug_tables exposes)
constexpr int AUX_ZG = 0;                       // [NZ]
constexpr int AUX_DL = AUX_ZG + NZ;             // [NZ]
constexpr int AUX_DDL = AUX_DL + NZ;            // [NZ]
constexpr int AUX_DVC = AUX_DDL + NZ;           // [NZ]
constexpr int AUX_TAN = AUX_DVC + NZ;           // [3 tables: dl, ddl, dvc][3 params: Om, w, wa][NZ]
constexpr int AUX_G = AUX_TAN + 9 * NZ;         // [6][NM]: log_dN_grid, d/d(a, b, mpisn, mbhmax, sigma)
constexpr int AUX_DOUBLES = AUX_G + 6 * NM;

constexpr int PRO_THREADS = 256;

// Read-only view that always loads through L2 (the data were produced by other CTAs of the same launch).
struct CgView {
    const double* p;
    __device__ __forceinline__ double operator[](int i) const { return __ldcg(p + i); }
    __device__ __forceinline__ CgView operator+(int o) const { return CgView{p + o}; }
};

struct EvalConsts {   // theta-independent numbers the prologue copies into the scalar block
    double log_nsamp;
    double log_ndraw;
    int use_wa;
    int fixed;
};

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /* >= 8*NV doubles */) {
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
        for (int w = 0; w < PRO_THREADS / 32; ++w) s += red[w * NV + k];
        v[k] = s;
    }
}

__device__ __forceinline__ double block_max(double x, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    double m = red[0];
    for (int w = 1; w < PRO_THREADS / 32; ++w) m = fmax(m, red[w]);
    return m;
}

// ---------------------------------------------------------------- PISN row (Dual<5>: a, b, mpisn, mbhmax, sigma)
__device__ void pisn_row(const double* __restrict__ th, int i, double* __restrict__ aux, double* sm) {
    typedef Dual<5> D;
    const int j = threadIdx.x;
    D a = D::var(th[T_A], 0), b = D::var(th[T_B], 1), mpisn = D::var(th[T_MPISN], 2),
      M = D::var(th[T_MBHMAX], 3), sg = D::var(th[T_SIGMA], 4);
    D top = M + 7.0 * sg;                                 // :99
    D mcomax = 2.0 * M - mpisn;                           // :29
    D mco_top = mcomax + dsqrt(4.0 * M * (M - mpisn));    // :30
    const double si = (double)i / (NM - 1), sj = (double)j / (NM - 1);
    D mbh = (i == NM - 1) ? top : (MIN_BH_MASS * (1.0 - si) + top * si);        // :102 linspace
    D mco = (j == NM - 1) ? mco_top : (MIN_CO_MASS * (1.0 - sj) + mco_top * sj); // :103
    D alpha = 1.0 / (4.0 * (mpisn - M));                  // :22
    D mu = (mco.v < mpisn.v) ? mco : (M + alpha * dsquare(mco - mcomax));        // :25
    D lx = dlog(mco / MTR);
    D ell = (mco.v < MTR) ? (-a * lx) : (-b * lx);        // :43
    D u = (mbh - mu) / sg;
    D lw = ell - 0.5 * dsquare(u) - HALF_LOG_2PI - dlog(sg);   // :105

    // neighbour exchange through shared memory: sm[0..6*NM) = lw (v, d[5]); sm[6*NM..) = mco (v, d_mpisn, d_mbhmax)
    double* s_lw = sm;
    double* s_mco = sm + 6 * NM;
    s_lw[j] = lw.v;
#pragma unroll
    for (int k = 0; k < 5; ++k) s_lw[(k + 1) * NM + j] = lw.d[k];
    s_mco[j] = mco.v;
    s_mco[NM + j] = mco.d[2];
    s_mco[2 * NM + j] = mco.d[3];
    __syncthreads();
    D term(-INFINITY);
    if (j < NM - 1) {
        D lw1;
        lw1.v = s_lw[j + 1];
#pragma unroll
        for (int k = 0; k < 5; ++k) lw1.d[k] = s_lw[(k + 1) * NM + j + 1];
        D dm(s_mco[j + 1] - mco.v);
        dm.d[2] = s_mco[NM + j + 1] - mco.d[2];
        dm.d[3] = s_mco[2 * NM + j + 1] - mco.d[3];
        term = dlogaddexp(lw1, lw) + dlog(dm) - LN2;      // :106  log(0.5) + logaddexp + log(diff)
    }
    double* red = sm + 9 * NM;
    const double mx = block_max(term.v, red);             // :107  logsumexp over the 255 cells
    double acc[6];
    const double e = (j < NM - 1) ? exp(term.v - mx) : 0.0;
    acc[0] = e;
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k + 1] = (j < NM - 1) ? e * term.d[k] : 0.0;
    block_sum<6>(acc, red);
    if (j == 0) {
        aux[AUX_G + i] = mx + log(acc[0]);
#pragma unroll
        for (int k = 0; k < 5; ++k) aux[AUX_G + (k + 1) * NM + i] = acc[k + 1] / acc[0];
    }
}

// ---------------------------------------------------------------- cosmology tables (Dual<3>: Om, w, wa)
// The 1024 knots are split over COS_CHUNKS blocks (one knot per thread): a single block was the critical path of the
// whole prologue.  The four chunks are the four blocks of ONE thread-block cluster (co-scheduled by the hardware):
// the cumulative trapezoid crosses chunks through distributed shared memory - every block publishes its chunk total in
// its own shared memory, one cluster barrier, then each block adds the totals of the chunks before it.
constexpr int COS_CHUNKS = 4;

constexpr int COS_VALS = 13;        // z, dl, ddl, dvc and the 9 tangents of one knot
constexpr int COS_EX0 = 80;         // offset of the neighbour-exchange area in the block's shared array
constexpr int PRO_SMEM_DOUBLES = COS_EX0 + COS_VALS * PRO_THREADS;
static_assert(PRO_SMEM_DOUBLES >= 9 * NM + 64, "the PISN rows use 9 NM + 64 doubles of the same array");

// One bin of the packed cosmology tables: lo / hi = the 13 numbers of its left / right knot (both the last knot for the
// beyond-the-table record NZ-1: all slopes zero), own = those of knot b itself.  Formats: bump_layout.cuh.
__device__ __forceinline__ void pack_cosmology_bin(const int b, const double* lo, const double* hi, const double* own,
                                                   double* __restrict__ blob, const EvalConsts ec, int& j0, int& j1) {
    j0 = j1 = 0;   // this bin's range of the d_L bucket table (empty in fixed-cosmology mode and for the padding bin)
    double2* cos = reinterpret_cast<double2*>(blob + OFF_COS);
    const int b0 = min(b, NZ - 2);
    cos[CR_DL * NZ + b] = make_double2(lo[1], b <= NZ - 2 ? 1.0 / (hi[1] - lo[1]) : 0.0);
    cos[CR_DVC * NZ + b] = make_double2(lo[3], hi[3] - lo[3]);
    cos[CR_DDL * NZ + b] = make_double2(lo[2], hi[2] - lo[2]);
    cos[CR_Z * NZ + b] = make_double2(1.0 / (1.0 + lo[0]), b <= NZ - 2 ? b0 * ZSTEP : LOG_ZMAX1);
    if (ec.fixed) return;
    // tangent tables: value order [dl, ddl, dvc][Om, w, wa] -> blob order CosTan; knot values in w0-wa mode, per-bin
    // pairs {t_b, t_{b+1} - t_b} otherwise, nothing in fixed-cosmology mode
    const int dst[9] = {CT_DL_OM, CT_DL_W, CT_DL_WA, CT_DDL_OM, CT_DDL_W, CT_DDL_WA, CT_DVC_OM, CT_DVC_W, CT_DVC_WA};
    double* ctan = blob + OFF_CTAN;
    if (ec.use_wa) {
#pragma unroll
        for (int r = 0; r < 9; ++r) ctan[dst[r] * NZ + b] = own[4 + r];
    } else {
        double2* ctan2 = reinterpret_cast<double2*>(ctan);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            if (r % 3 == 2) continue;   // the wa tangents
            ctan2[dst[r] * NZ + b] = make_double2(lo[4 + r], hi[4 + r] - lo[4 + r]);
        }
    }
    // ---- bucket table of the d_L search: srch[j] = a bin index that is <= the bin of every x in bucket j, i.e.
    // for the smallest double x0_j of the bucket: clip(#{k < NZ-1 : dl_k <= x0_j}, 1, .) - 1.  Knot b owns the
    // buckets whose x0_j lies in [dl_b, dl_{b+1}) (the last knot up to +inf): no search.  The caller fills the ranges
    // warp-cooperatively (the first knots own hundreds of buckets each: d_L grows linearly from 0 there).
    if (b <= NZ - 2) {
        auto first_bucket_at_or_above = [](const double v) -> int {   // min{j : x0_j >= v}, clamped to [0, SRCH_N]
            if (!(v > 0.0)) return 0;
            const int key = (__double2hiint(v) >> (20 - SRCH_MBITS)) - (SRCH_EXP_LO << SRCH_MBITS);
            if (key < 0) return 0;
            if (key >= SRCH_N) return SRCH_N;
            const bool exact = __double2loint(v) == 0 && (__double2hiint(v) & ((1 << (20 - SRCH_MBITS)) - 1)) == 0;
            return exact ? key : key + 1;
        };
        j0 = (b == 0) ? 0 : first_bucket_at_or_above(lo[1]);
        j1 = (b == NZ - 2) ? SRCH_N : first_bucket_at_or_above(hi[1]);
    }
}

// Where the prologue puts a scalar: into the blob's scalar block (read by the epilogue and by the debugging entry
// points) and, unless the build takes the stream kernel's scalars from the blob, straight into the context's slot of
// the constant bank through the slot's global address.  A kernel may not write constant memory that IT reads; the
// prologue never reads K_SC4, and the stream kernel that does is a later launch.  (Round 1 copied the 512 bytes with
// a device-to-device copy node between the two kernels: ~1.1 us of every evaluation.)
struct ScalarSink {
    double* blob_scal;
    double* cbank;   // may be null
    __device__ __forceinline__ void set(const int i, const double v) const {
        blob_scal[i] = v;
        if (cbank) cbank[i] = v;
    }
};

// Returns non-zero if one of this thread's table values is not finite (theta outside the prior support).
__device__ int cosmology_tables(const double* __restrict__ th, const EvalConsts ec, double* __restrict__ aux,
                                double* __restrict__ blob, const ScalarSink scal, double* sm, const int chunk) {
    const int use_wa = ec.use_wa;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    typedef Dual<3> D;
    const int tid = chunk * PRO_THREADS + threadIdx.x;   // global knot-thread index
    const double h = th[T_H];
    D Om = D::var(th[T_OM], 0), w = D::var(th[T_W], 1), wa = D::var(use_wa ? th[T_WA] : 0.0, 2);
    const double dH = C_H100_GPC / h;   // :239
    constexpr int PER = NZ / (PRO_THREADS * COS_CHUNKS);  // 1 knot per thread (+ its right neighbour, recomputed)
    double z[PER + 1];
    D iE[PER + 1];
#pragma unroll
    for (int q = 0; q <= PER; ++q) {
        const int k = tid * PER + q;
        const double lz = (k >= NZ - 1) ? LOG_ZMAX1 : k * ZSTEP;   // np.linspace(0, log(101), 1024), :230
        z[q] = expm1(lz);
        const double opz = 1.0 + z[q];
        const double lopz = lz;   // log(1 + expm1(lz)): the grid IS uniform in log(1+z) (saves a libm log on the chain)
        D de = dexp((3.0 * (1.0 + w + wa)) * lopz);               // opz**(3(1+w)), :256
        if (use_wa) de = de * dexp(wa * (-3.0 * z[q] / opz));     // CPL extension (no reference counterpart)
        D E = dsqrt(Om * (opz * opz * opz) + (1.0 - Om) * de);
        iE[q] = 1.0 / E;
    }
    // per-thread increments of the cumulative trapezoid (utils.py:8), then a block-wide exclusive scan
    D inc[PER];
    D tot(0.0);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int k = tid * PER + q;
        inc[q] = (k < NZ - 1) ? (0.5 * (z[q + 1] - z[q])) * (iE[q] + iE[q + 1]) : D(0.0);
        tot = tot + inc[q];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sc[4] = {tot.v, tot.d[0], tot.d[1], tot.d[2]};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double n = __shfl_up_sync(0xffffffffu, sc[c], o);
            if (lane >= o) sc[c] += n;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sm[warp * 4 + c] = sc[c];
    }
    __syncthreads();
    // chain across chunks: publish this chunk's total in this block's shared memory (sm[72..75]) ...
    if (threadIdx.x == PRO_THREADS - 1) {
        // sc[] of the last thread of the last warp is that warp's inclusive total; add the earlier warps
        double t4[4] = {sc[0], sc[1], sc[2], sc[3]};
        for (int ww = 0; ww < PRO_THREADS / 32 - 1; ++ww) {
#pragma unroll
            for (int c = 0; c < 4; ++c) t4[c] += sm[ww * 4 + c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sm[72 + c] = t4[c];
    }
    cluster.sync();   // ... every chunk's total is now visible cluster-wide (release / acquire)
    if (threadIdx.x == 0) {
        double p4[4] = {0, 0, 0, 0};
        for (int cc = 0; cc < chunk; ++cc) {   // chunk == rank of this block in its cluster (blocks 0..3 of the grid)
            const double* peer = cluster.map_shared_rank(sm + 72, cc);
#pragma unroll
            for (int c = 0; c < 4; ++c) p4[c] += peer[c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sm[64 + c] = p4[c];
    }
    __syncthreads();  // sm[64..67]; the totals' slots are not reused, and no block exits before the barriers below
    double base[4] = {sm[64], sm[65], sm[66], sm[67]};
    for (int ww = 0; ww < warp; ++ww) {
#pragma unroll
        for (int c = 0; c < 4; ++c) base[c] += sm[ww * 4 + c];
    }
    static_assert(PER == 1, "one knot per thread: the packing below exchanges one value set with the right neighbour");
    D C;  // exclusive prefix for this thread's knot
    C.v = base[0] + sc[0] - tot.v;
#pragma unroll
    for (int c = 0; c < 3; ++c) C.d[c] = base[c + 1] + sc[c + 1] - tot.d[c];
    const int k = tid;
    const double opz = 1.0 + z[0];
    const D dc = dH * C;                                        // :231
    const D dl = dc * opz;                                      // :232
    const D ddl = dc + (dH * opz) * iE[0];                      // :233
    const D dvc = (FOUR_PI * dH) * (dsquare(dc) * iE[0]);       // :235
    // the knot's 13 numbers in aux order: z, dl, ddl, dvc, then the tangents [dl, ddl, dvc][Om, w, wa]
    double me[COS_VALS] = {z[0], dl.v, ddl.v, dvc.v, dl.d[0], dl.d[1], dl.d[2], ddl.d[0], ddl.d[1], ddl.d[2],
                           dvc.d[0], dvc.d[1], dvc.d[2]};
#pragma unroll
    for (int r = 0; r < COS_VALS; ++r) aux[r * NZ + k] = me[r];   // raw knots (bump_debug_tables)
    if (k == 0) scal.set(S_DL_FIRST, dl.v);
    if (k == NZ - 1) scal.set(S_DL_LAST, dl.v);
    static_assert(AUX_ZG == 0 && AUX_DL == NZ && AUX_DDL == 2 * NZ && AUX_DVC == 3 * NZ && AUX_TAN == 4 * NZ,
                  "aux order of the 13 cosmology value sets");
    // ---- packed per-bin records, straight from the registers: bin b = [knot b, knot b+1] needs the right neighbour's
    // numbers - the next thread's through shared memory, the next chunk's first thread's through distributed shared
    // memory.  (Round 1 packed them in a second kernel, round 2 first in the prologue's last block: ~7 us of its tail.)
    double* ex = sm + COS_EX0;   // [COS_VALS][PRO_THREADS]
#pragma unroll
    for (int r = 0; r < COS_VALS; ++r) ex[r * PRO_THREADS + threadIdx.x] = me[r];
    cluster.sync();
    double lo[COS_VALS], hi[COS_VALS];
    // record NZ-1 is the bin BEYOND the table: the last knot's values with zero slopes, which is what jnp.interp returns
    // for x > xp[-1] (fp[-1], no gradient through x) - the streaming kernel then needs no special case for such samples
    // (w0-wa mode keeps knot-value tangent tables, which have no slot for it: that kernel clamps and selects instead)
    const bool last_knot = (k == NZ - 1);
    if (!last_knot) {
        const double* src = (threadIdx.x + 1 < PRO_THREADS) ? ex + threadIdx.x + 1
                                                            : cluster.map_shared_rank(ex, chunk + 1);
#pragma unroll
        for (int r = 0; r < COS_VALS; ++r) lo[r] = me[r], hi[r] = src[r * PRO_THREADS];
    } else {
#pragma unroll
        for (int r = 0; r < COS_VALS; ++r) lo[r] = me[r], hi[r] = me[r];
    }
    int j0, j1;
    pack_cosmology_bin(k, lo, hi, me, blob, ec, j0, j1);
    {
        unsigned short* srch = reinterpret_cast<unsigned short*>(blob + OFF_SRCH);
        // short ranges (almost all: a bin is a handful of buckets wide) by their own thread, long ones (the first
        // knots: hundreds of buckets each) by all lanes of the warp
        constexpr int LONG_RANGE = 16;
        const bool is_long = j1 - j0 > LONG_RANGE;
        if (!is_long)
            for (int j = j0; j < j1; ++j) srch[j] = (unsigned short)(j == 0 ? 0 : k);
        unsigned int todo = __ballot_sync(0xffffffffu, is_long);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int a0 = __shfl_sync(0xffffffffu, j0, src), a1 = __shfl_sync(0xffffffffu, j1, src);
            const int bb = __shfl_sync(0xffffffffu, k, src);
            for (int j = a0 + lane; j < a1; j += 32) srch[j] = (unsigned short)(j == 0 ? 0 : bb);
        }
    }
    cluster.sync();   // no block may exit (and release its shared memory) while a peer still reads its knot
    int bad = 0;
#pragma unroll
    for (int r = 1; r < COS_VALS; ++r) bad |= !isfinite(me[r]);
    // The streaming kernel takes a single step from srch[j]: a bucket is narrower than any bin, so that is exact unless
    // one of the two clamped end buckets spans more than two bins (first: every x below 2^-8 (1 + 1/256) Gpc must lie in
    // bin 0 or 1; last: every x above 2^13 (1 - 1/256) Gpc in one of the last two bins).  That takes roughly h > 7 or
    // h < 0.11 (prior: 0.35 .. 1.4): flag it (-> NaN outputs).  Checked by the two threads that own those knots.
    if (!ec.fixed) {
        const double first_hi = __hiloint2double(((SRCH_EXP_LO << SRCH_MBITS) + 1) << (20 - SRCH_MBITS), 0);
        const double last_lo = __hiloint2double(((SRCH_EXP_LO << SRCH_MBITS) + SRCH_N - 1) << (20 - SRCH_MBITS), 0);
        if (k == 2) bad |= !(dl.v >= first_hi);
        if (k == NZ - 3) bad |= !(dl.v <= last_lo);
    }
    return bad;
}

// ---------------------------------------------------------------- scalars
// Called by ALL 32 lanes of one warp.  log_pl_norm (:136) and log_norm (:138, :140-151) with their tangents are seven
// forward-mode chains, one per variable v of (a, b, c, mpisn, mbhmax, sigma, fpl) (a Dual<1> each instead of one thread
// with a Dual<7>), and every chain is split over three lanes, lane = 3 v + sub, that evaluate its independent pieces
//   sub 0: log_pl_norm = log fpl + PISN(mbhmax)      sub 1: PISN(mref)      sub 2: turn-on(mref) - c log(mref / mbhmax)
// before lane 3 v joins them with two shuffles (this is the longest serial piece of the whole prologue: ~1000 dependent
// FP64 instructions per chain unsplit).  Lane 21 derives the rate normalisation (-self(zref = 0), :168,173), lane 22
// the theta-only numbers.  gtab = [6][NM] copy of aux[AUX_G ...] in shared memory.
__device__ void build_scalars(const double* th, const double* gtab, const EvalConsts ec, const ScalarSink scal,
                              const int lane) {
    typedef Dual<1> D;
    const int v = lane / 3, sub = lane - 3 * v;
    D piece(0.0);        // this lane's piece of chain v
    double lnV = 0.0;    // lane 21
    if (lane < 21) {
        const int map5[5] = {0, 1, 3, 4, 5};   // (a, b, mpisn, mbhmax, sigma) -> variable index
        int q5 = -1;                           // this chain's row in the PISN tangent table, if any
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (map5[q] == v) q5 = q;
        auto seed = [&](const double x, const int var) {
            D r(x);
            r.d[0] = (var == v) ? 1.0 : 0.0;
            return r;
        };
        const D c = seed(th[T_C], 2), M = seed(th[T_MBHMAX], 4), sg = seed(th[T_SIGMA], 5), fpl = seed(th[T_FPL], 6);
        const D top = M + 7.0 * sg;
        auto knot = [&](int k) -> D {
            const double s = (double)k / (NM - 1);
            return (k == NM - 1) ? top : (MIN_BH_MASS * (1.0 - s) + top * s);
        };
        auto Gk = [&](int k) -> D {
            D g(gtab[k]);
            g.d[0] = (q5 >= 0) ? gtab[(q5 + 1) * NM + k] : 0.0;
            return g;
        };
        // jnp.interp(m, mbh_grid, log_dN_grid), differentiable in m, the knots and the values (:110-111)
        auto pisn = [&](const D& m) -> D {
            int i = (int)floor((m.v - MIN_BH_MASS) / (top.v - MIN_BH_MASS) * (NM - 1)) + 1;
            i = min(max(i, 1), NM - 1);
            while (i < NM - 1 && knot(i).v <= m.v) ++i;       // i = clip(#{knots <= m}, 1, n-1)
            while (i > 1 && knot(i - 1).v > m.v) --i;
            D x0 = knot(i - 1), x1 = knot(i), f0 = Gk(i - 1), f1 = Gk(i);
            D f = f0 + ((m - x0) / (x1 - x0)) * (f1 - f0);
            if (m.v < knot(0).v) f = Gk(0);
            if (m.v > top.v) f = Gk(NM - 1);
            return f;
        };
        const D mref(MREF);
        if (sub == 0) {
            piece = dlog(fpl) + pisn(M);                                                        // log_pl_norm, :136
        } else if (sub == 1) {
            piece = (MREF <= MIN_BH_MASS || MREF >= top.v) ? D(-INFINITY) : pisn(mref);         // :144-145
        } else {
            piece = LN2 - dlog1p(dexp(-(mref - M) / (M * TURNON_WIDTH))) - c * dlog(mref / M);  // :52-54, :147
        }
    } else if (lane == 21) {
        const double kappa = th[T_KAPPA], zp = th[T_ZP];
        const double lopzp = log1p(zp);
        const double r0 = exp(-kappa * lopzp);
        lnV = log1p(r0);
        const double sig0 = r0 / (1.0 + r0);
        scal.set(S_KAPPA, kappa);
        scal.set(S_ZP, zp);
        scal.set(S_LOPZP, lopzp);
        scal.set(S_RATE_LOG_NORM, lnV);
        scal.set(S_LNV_KAPPA, -sig0 * lopzp);
        scal.set(S_LNV_ZP, -sig0 * kappa / (1.0 + zp));
    } else if (lane == 22) {
        const double M = th[T_MBHMAX], top = M + 7.0 * th[T_SIGMA];
        scal.set(S_H, th[T_H]);
        scal.set(S_INV_H, 1.0 / th[T_H]);
        scal.set(S_C, th[T_C]);
        scal.set(S_M, M);
        scal.set(S_LOG_M, log(M));
        scal.set(S_INV_DM, 1.0 / (M * TURNON_WIDTH));
        scal.set(S_TOP, top);
        const double inv_dmbh = (NM - 1) / (top - MIN_BH_MASS);
        scal.set(S_INV_DMBH, inv_dmbh);
        scal.set(S_INV_TOPM3, 1.0 / (top - MIN_BH_MASS));
        scal.set(S_BETA, th[T_BETA]);
        scal.set(S_LAM, th[T_LAM]);
        scal.set(S_FPL, th[T_FPL]);
        scal.set(S_LOG_NSAMP, ec.log_nsamp);
        scal.set(S_LOG_NDRAW, ec.log_ndraw);
        scal.set(S_USE_WA, (double)ec.use_wa);
        scal.set(S_FIXED, (double)ec.fixed);
        scal.set(S_ZEPS, expm1(ZSTEP));
        scal.set(S_POS0, -MIN_BH_MASS * inv_dmbh);
        scal.set(S_LAM2, th[T_LAM] - 2.0);
        scal.set(S_RATE0, th[T_LAM] - 3.0 - th[T_BETA]);
    }
    __syncwarp();
    // join: lane 3 v (sub 0) takes PISN(mref) from lane 3 v + 1 and the power-law piece from lane 3 v + 2
    D P, T;
    P.v = __shfl_down_sync(0xffffffffu, piece.v, 1), P.d[0] = __shfl_down_sync(0xffffffffu, piece.d[0], 1);
    T.v = __shfl_down_sync(0xffffffffu, piece.v, 2), T.d[0] = __shfl_down_sync(0xffffffffu, piece.d[0], 2);
    lnV = __shfl_sync(0xffffffffu, lnV, 21);
    if (lane >= 21 || sub != 0) return;
    const D lpn = piece;
    const D Q = T + lpn;                                                   // -c log(m/mbhmax) + log_pl_norm + turn-on
    const D A0 = (MREF < MBH_MIN) ? D(-INFINITY) : dlogaddexp(P, Q);       // :149-150
    const D ln = -(A0 + log(MREF));                                        // :138
    {
        const int map5[5] = {0, 1, 3, 4, 5};
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (map5[q] == v) scal.set(S_LPN_D0 + q, lpn.d[0]);
    }
    scal.set(S_LN_D0 + v, ln.d[0]);
    if (v != 0) return;
    scal.set(S_LPN, lpn.v);
    scal.set(S_LOG_NORM, ln.v);
    scal.set(S_EXP_LPN, exp(lpn.v));
    scal.set(S_C2, 2.0 * exp(lpn.v));
    scal.set(S_LOG_C2, LN2 + lpn.v);
    scal.set(S_CONST, 2.0 * ln.v + lnV - th[T_BETA] * LOG_MREF_PAIR);
}

// ---------------------------------------------------------------- packed mass records (last block of the prologue)
// Bin i of the mbh grid needs rows i and i + 1 of the PISN table, i.e. two different blocks: packed once all rows are
// done, by warps 1..7 of the last block while warp 0 derives the scalars.  aux was written by other blocks: read via L2.
__device__ void pack_mass_records(const double* __restrict__ gtab /* shared copy [6][NM] */, double* __restrict__ blob,
                                  const int t /* 0 .. nt-1 */, const int nt) {
    double2* mass = reinterpret_cast<double2*>(blob + OFF_MASS);
    for (int item = t; item < NMREC * NM; item += nt) {
        const int r = item / NM, i = item - r * NM;
        if (i <= NM - 2) {
            const double g0 = gtab[r * NM + i];
            mass[item] = make_double2(g0, gtab[r * NM + i + 1] - g0);
        } else {
            // record NM-1 is the bin BEYOND the grid (m >= mbhmax + 7 sigma): log dN = -inf there (:145).  A value of
            // -5e4 makes the kernel's exp return its saturation value (~1e-304 relative to the power-law term: below
            // every rounding), with zero slope and zero tangents - no compare-and-select per sample in the kernel.
            mass[item] = make_double2(r == MR_G ? MASS_BEYOND_LOG : 0.0, 0.0);
        }
    }
}

// ---------------------------------------------------------------- the prologue kernel
//   blocks 0..3        : flat wCDM distance tables, 256 knots each (intensity_models.py:229-235 + utils.py:3-8); they
//                        are one cluster and chain their cumulative trapezoid through distributed shared memory
//   blocks 4..4+NM-1   : row i of the PISN pile-up table  (intensity_models.py:96-108, LogDNDMPISN.__post_init__)
//                        and pack their bins of the cosmology records and of the d_L bucket table themselves
//   the row block that finishes last (rows ticket): packed mass records (warps 1..7) and the scalars
//                        (intensity_models.py:134-138,167-168; 9 threads of warp 0) -> the blob the streaming kernel stages
//   the block that finishes last of all (ticket): the validity flag
constexpr int PRO_BLOCKS = NM + COS_CHUNKS;
static_assert(PRO_BLOCKS % COS_CHUNKS == 0, "the grid is a whole number of clusters");

__global__ void __cluster_dims__(COS_CHUNKS, 1, 1) __launch_bounds__(PRO_THREADS)
prologue_kernel(const double* __restrict__ theta, double* __restrict__ aux, double* __restrict__ blob,
                double* __restrict__ cbank /* global address of the context's constant-bank slot, or null */,
                unsigned int* __restrict__ flags /* [0] ticket (low half) + bad count (high half), [2] rows ticket */,
                const EvalConsts ec, unsigned long long* __restrict__ tl) {
    pdl_launch_dependents<PDL_STREAM && BUMP_PDL_STREAM_TRIGGER == 2>();   // (measurement option: at kernel start)
    const ScalarSink scal = {blob + OFF_SCAL, cbank};
    __shared__ double sm[PRO_SMEM_DOUBLES];
    __shared__ double th[NTHETA_MAX];
    __shared__ bool is_last;
    timeline_begin(tl, TL_PROLOGUE);
    const int use_wa = ec.use_wa;
    if (threadIdx.x < NTHETA_MAX) th[threadIdx.x] = (threadIdx.x < NTHETA || use_wa) ? theta[threadIdx.x] : 0.0;
    __syncthreads();
    const bool is_cos = blockIdx.x < COS_CHUNKS;    // cosmology chunks first: they are the longer chains
    const int row = (int)blockIdx.x - COS_CHUNKS;
    timeline_begin(tl, is_cos ? TL_PRO_COSMO : TL_PRO_ROWS);
    int bad = 0;
    if (!is_cos) pisn_row(th, row, aux, sm);
    else bad = cosmology_tables(th, ec, aux, blob, scal, sm, blockIdx.x);
    timeline_end(tl, is_cos ? TL_PRO_COSMO : TL_PRO_ROWS);
    // non-finite tables (theta outside the prior support) or theta: flag it, finalize returns NaN.
    // (the PISN table is checked by the row block that finishes last, below)
    // The verdict travels in the upper half of the ticket word (every block adds 1 + 0x10000 if it saw something bad),
    // so that the block that finishes last knows it from the value its own ticket returns: no second round trip.
    __shared__ unsigned int s_bad;
    if (threadIdx.x == 0) s_bad = 0u;
    if (is_cos) {
        if (threadIdx.x < NTHETA_MAX) bad |= !isfinite(th[threadIdx.x]);
        const int any_bad = __syncthreads_or(bad);
        if (threadIdx.x == 0) s_bad = any_bad ? 1u : 0u;
    }
    // ---- the LAST ROW block (rows ticket): once all 256 PISN rows exist, the scalars (seven forward-mode chains of
    // ~1000 dependent FP64 instructions each: the longest serial piece of the prologue) and the packed mass records.
    // It does not wait for the cosmology cluster, which is still at work then (rows: ~5 us, cosmology: ~9.5 us).
    if (!is_cos) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = (atomicAdd(flags + 2, 1u) == NM - 1);
        __syncthreads();
        if (is_last) {
            __threadfence();
            timeline_begin(tl, TL_PRO_LAST);
            double* gtab = sm;   // [6][NM] copy of the PISN table for the scalars
            int bad_g = 0;   // a non-finite PISN table (theta outside the prior support): flag it, finalize returns NaN
            {   // all six loads of a thread in flight together: consumed one by one they cost six L2 round trips in a row
                static_assert(6 * NM == 6 * PRO_THREADS, "six table values per thread");
                double g[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) g[r] = __ldcg(aux + AUX_G + r * PRO_THREADS + threadIdx.x);
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    gtab[r * PRO_THREADS + threadIdx.x] = g[r];
                    bad_g |= !isfinite(g[r]);
                }
            }
            {
                const int any_bad = __syncthreads_or(bad_g);
                if (threadIdx.x == 0 && any_bad) s_bad = 1u;
            }
            if (threadIdx.x < 32) {
                build_scalars(th, gtab, ec, scal, threadIdx.x);
                if (threadIdx.x == 0) flags[2] = 0u;   // re-arm the rows ticket
            } else {
                pack_mass_records(gtab, blob, threadIdx.x - 32, PRO_THREADS - 32);
            }
            __syncthreads();
            timeline_end(tl, TL_PRO_LAST);
        }
    }
    // ---- the block that finishes last of all (global ticket): the validity flag, and re-arming
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int mine = 1u + (s_bad ? 0x10000u : 0u);
        const unsigned int old = atomicAdd(flags, mine);
        if ((old & 0xffffu) == gridDim.x - 1) {
            scal.set(S_BAD, ((old + mine) >> 16) != 0u ? 1.0 : 0.0);
            flags[0] = 0u;   // re-arm the ticket for the next evaluation
        }
    }
    timeline_end(tl, TL_PROLOGUE);
}

}  // namespace bump

// libbump_b200.so — C ABI (include/bump.h) over the sm_100a kernels.  CUDA runtime only; no torch, no CPU fallback.
//
// One evaluation = 3 kernel launches replayed from a CUDA graph, the second and third as programmatic dependents of
// their predecessors (their blocks are scheduled while the predecessor drains and wait in griddepcontrol.wait):
//   prologue_kernel  (bump_tables.cuh)   theta -> PISN and cosmology tables + tangents (F1-F2 of SURVEY.md 2.2), then
//                                        in its last block the packed per-bin records, d_L bucket table, scalars (F3;
//                                        the scalars also go straight into this context's constant-bank slot)
//   stream_kernel    (bump_stream.cuh)   one pass over the SoA columns -> per-warp records (F4-F7 + reverse pass)
//   epilogue_kernel  (bump_epilogue.cuh) per-event logsumexp / Neff, rank partial, and (single rank, or peer-memory
//                                        exchange) the result
//   [NCCL exchange: ncclAllGather of the 1 KiB partial, then finalize_kernel: rank-ordered merge -> result header]
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bump.h"
#include "bump_epilogue.cuh"
#include "bump_layout.cuh"
#include "bump_stream.cuh"
#include "bump_tables.cuh"

using namespace bump;

static_assert(BUMP_NTHETA == NTHETA && BUMP_NTHETA_MAX == NTHETA_MAX, "theta layout");
static_assert(BUMP_OUT_HEADER == OUT_HEADER && BUMP_OUT_DLOGLIKE == OUT_DLOGLIKE && BUMP_OUT_DLOG_MU == OUT_DLOG_MU,
              "output layout");
static_assert(BUMP_PARTIAL_LEN == PARTIAL_LEN, "partial layout");

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(BUMP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                         std::to_string(__LINE__) + ")");                                 \
    } while (0)

// ---- NCCL through dlopen: the library loads (and the single-GPU path runs) without libnccl
struct UniqueId {
    char internal[128];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return BUMP_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(BUMP_E_NCCL, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<int (*)(void**, int, UniqueId, int)>(dlsym(h, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(
        dlsym(h, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy)
        return fail(BUMP_E_NCCL, "libnccl is missing a required symbol");
    g_nccl = a;
    return BUMP_OK;
}
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble in nccl.h

#define NCK(call)                                                                                  \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != 0)                                                                               \
            return fail(BUMP_E_NCCL, std::string(#call) + ": " +                                   \
                                         (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
    } while (0)

// Host -> device copy that is complete, in the context's (non-blocking) stream order AND for the host, when it returns.
// A plain cudaMemcpy from pageable memory returns once the data is staged - the DMA into device memory may still be
// running - and it is ordered against the legacy default stream only, which a non-blocking stream does not wait for:
// a kernel launched on c->stream right afterwards could read the tail of the buffer before it arrived (seen once the
// reciprocal probe of bump_debug_math ran right behind its 2.4 MB input copy: NaN in the last few thousand entries).
#define H2D_SYNC(c, dst, src, bytes)                                                              \
    do {                                                                                          \
        CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, (c)->stream));          \
        CK(cudaStreamSynchronize((c)->stream));                                                   \
    } while (0)

// ---- upload-time kernel: raw (m1_det, q, d_L, pdraw) -> the 7 padded SoA columns (theta-independent logs hoisted;
// the reference recomputes log(pdraw) on every trace, intensity_models.py:365)

// Locality key of a sample.  Sorting the samples of an event (the likelihood is a sum over them, so their order is
// free) makes the 32 lanes of a warp land in the same or in neighbouring records of the d_L tables and of the mass
// table - at m1 AND at m2 = q m1: a shared-memory read of 16-byte records is conflict-free when the 8 lanes of a
// quarter-warp hit records that are equal or less than 8 apart (DESIGN.md "sample order").
//   key = row | d_L bucket (1/16 octave) | m1_det bucket (4 % wide) | m2_det (fine, 1/512 in log), the direction of
//   the last field alternating from one m1 bucket to the next so that the walk through the (m1, m2) plane is continuous.
// The buckets are in detector-frame quantities (theta-independent); inside one d_L bucket (1 + z) is the same for every
// sample to a few percent whatever the cosmology, so neighbours in (m1_det, m2_det) are neighbours in the source
// frame too.  Round 1's key (d_L 1/32 octave | m1_det fine) left m2 unordered: 6.7 wavefronts per LDS.128 at m2
// against a floor of 4 (ncu, O5 mock), 15 % of the kernel's shared-memory wavefronts (build option
// BUMP_SORT_KEY_M1_ONLY keeps it for comparison).
__global__ void locality_keys_kernel(const double* __restrict__ m1d, const double* __restrict__ q,
                                     const double* __restrict__ dl, const int64_t ncols, const int64_t n,
                                     unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long row = (unsigned long long)(i / ncols);   // < 2^27 (checked by the caller)
#ifdef BUMP_SORT_KEY_M1_ONLY
        const unsigned int kd = ((unsigned int)__double2hiint(dl[i]) >> 15) & 0xFFFFu;        // 1/32-octave buckets
        const unsigned int km = (unsigned int)min(max((__double2hiint(m1d[i]) - (1023 << 20)) >> 4, 0), 0xFFFFF);
        keys[i] = (row << 36) | ((unsigned long long)kd << 20) | km;
#else
        // (non-finite or non-positive inputs are rejected by prepare_columns_kernel right after; here they only
        // have to produce SOME key)
#ifndef BUMP_SORT_DL_BITS
#define BUMP_SORT_DL_BITS 4        // mantissa bits of the d_L bucket: 1/16 octave
#endif
#ifndef BUMP_SORT_M1_PER_LOG
#define BUMP_SORT_M1_PER_LOG 25.0  // m1_det buckets per unit of log: 0.04 wide
#endif
        static_assert(BUMP_SORT_DL_BITS >= 0 && BUMP_SORT_DL_BITS <= 8, "key layout: 27 | 11 + bits | 10 | the rest");
        constexpr int DLB = 11 + BUMP_SORT_DL_BITS;           // bits of the d_L field
        const unsigned int kd = ((unsigned int)__double2hiint(dl[i]) >> (20 - BUMP_SORT_DL_BITS)) & ((1u << DLB) - 1u);
        const double lm1 = log(m1d[i]), lm2 = lm1 + log(q[i]);
        const int km = min(max((int)floor(lm1 * BUMP_SORT_M1_PER_LOG) + 512, 0), 1023);       // 10 bits
        constexpr int K2B = 64 - 27 - 10 - DLB;               // bits left for m2 (12 with the default d_L field)
        constexpr int K2MAX = (1 << K2B) - 1;
        int k2 = min(max((int)floor((lm2 + 2.0) * (double)(1 << (K2B - 3))), 0), K2MAX);      // 0.14 .. 400 Msun, clamped
        if (km & 1) k2 = K2MAX - k2;
        keys[i] = (row << 37) | ((unsigned long long)kd << (10 + K2B)) | ((unsigned long long)km << K2B) | (unsigned long long)k2;
#endif
        idx[i] = (unsigned int)i;
    }
}

__global__ void prepare_columns_kernel(const double* __restrict__ m1d, const double* __restrict__ q,
                                       const double* __restrict__ dl, const double* __restrict__ pd,
                                       const unsigned int* __restrict__ perm, const int64_t nrows,
                                       const int64_t ncols, const int64_t stride, double* __restrict__ out,
                                       unsigned int* __restrict__ bad, const double* __restrict__ fixed_tab) {
    const int64_t total = nrows * stride;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / stride, c = i - r * stride;
        bool sentinel = c >= ncols;
        if (!sentinel) {
            int64_t s = r * ncols + c;
            if (perm) s = perm[s];
            const double vm = m1d[s], vq = q[s];
            // the kernels assume finite, positive inputs (the reference would produce NaN / -inf weights)
            if (!(vm > 0.0 && vq > 0.0 && dl[s] > 0.0 && pd[s] > 0.0 && isfinite(vm) && isfinite(vq) &&
                  isfinite(dl[s]) && isfinite(pd[s])))
                atomicOr(bad, 1u);
            // A sample whose DETECTOR-frame secondary is already below mbh_min has m2 = q m1_det / (1+z) < 5 at every
            // redshift: its weight is zero for every theta (intensity_models.py:149).  Storing a sentinel instead lets
            // the streaming kernel drop all range guards: what it sees always has m1 >= m2 >= 5/101.
            if (vq * vm < MBH_MIN || vm < MBH_MIN) sentinel = true;
            out[block_index(i, C_M1D)] = vm;
            out[block_index(i, C_Q)] = vq;
            out[block_index(i, C_LM)] = log(vm);
            out[block_index(i, C_LQ)] = log(vq);
            out[block_index(i, C_L1Q)] = log1p(vq);
            if (!fixed_tab) {
                out[block_index(i, C_DL)] = dl[s];
                out[block_index(i, C_LPD)] = log(pd[s]);
            } else {
                // fixed cosmology (pop_model, intensity_models.py:313-355): `dl` holds z; column 0 becomes log1p(z)
                // and log dVdzdt(z) = log interp(z, zinterp, dVdzdt_interp) (:332) is folded into the pdraw column
                const double z = dl[s], lz = log1p(z);
                int k = min(max((int)floor(lz / ZSTEP), 0), NZ - 2);
                while (k < NZ - 2 && expm1((k + 1) * ZSTEP) <= z) ++k;    // searchsorted(side='right') on zinterp
                while (k > 0 && expm1(k * ZSTEP) > z) --k;
                const double z0 = expm1(k * ZSTEP), z1 = (k + 1 == NZ - 1) ? expm1(LOG_ZMAX1) : expm1((k + 1) * ZSTEP);
                double v = fixed_tab[k] + (z - z0) / (z1 - z0) * (fixed_tab[k + 1] - fixed_tab[k]);
                if (z > z1 && k == NZ - 2) v = fixed_tab[NZ - 1];
                out[block_index(i, C_DL)] = lz;
                out[block_index(i, C_LPD)] = log(pd[s]) - log(v);
                if (!(v > 0.0)) sentinel = true;   // log dVdzdt = -inf: zero weight
            }
        }
        if (sentinel) {   // padding / never-valid sample: source mass 1/(1+z) < mbh_min -> weight exactly zero (-inf log weight)
            out[block_index(i, C_DL)] = 1.0;
            out[block_index(i, C_M1D)] = 1.0;
            out[block_index(i, C_Q)] = 1.0;
            out[block_index(i, C_LM)] = 0.0;
            out[block_index(i, C_LQ)] = 0.0;
            out[block_index(i, C_L1Q)] = LN2;
            out[block_index(i, C_LPD)] = 0.0;
        }
    }
}

// Accuracy probe of the streaming kernel's scalar math (bump_debug_math): y[i] = f(x[i]) with the exp table staged
// exactly where the kernel expects it.
__global__ void math_probe_kernel(const int which, const double* __restrict__ x, const int64_t n,
                                  const double* __restrict__ g_blob, double* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char probe_smem[];
    double* s = reinterpret_cast<double*>(probe_smem);
    for (int k = threadIdx.x; k < EXPT_DOUBLES; k += blockDim.x) s[OFF_EXPT + k] = g_blob[OFF_EXPT + k];
    __syncthreads();
    const uint32_t sb = smem_u32(probe_smem);
    const uint32_t rep = (uint32_t)(threadIdx.x & (EXPT_REPL - 1)) << 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        double r = 0.0;
        switch (which) {
            case 0: r = fexp<false>(v, sb, rep); break;
            case 1: r = fexp<true>(v, sb, rep); break;
            case 2: r = frcp(v); break;
            case 3: r = flog1p_small(v); break;
            default: r = frcp1p_small(v); break;
        }
        y[i] = r;
    }
}

struct DataSet {
    int64_t nrows = 0, ncols = 0, stride = 0;   // events: [nobs, nsamp]; injections: [1, nsel]
    double* base = nullptr;                     // nrows * stride / GROUP blocks of NCOL * GROUP doubles
    // The resident columns are read-only after upload and may be shared by several contexts (bump_ctx_clone: one
    // upload, one copy in HBM, one context per chain): freed by whoever drops the last reference.
    std::shared_ptr<void> owner;
    void reset() { *this = DataSet(); }
    int alloc(size_t bytes) {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return 1;
        base = static_cast<double*>(p);
        owner = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
        return 0;
    }
};

}  // namespace

struct bump_ctx {
    int device = 0;
    int slot = 0;                     // constant-bank slot of the streaming kernel's scalars (0 .. NSLOT-1)
    double* d_cbank = nullptr;        // global address of that slot (the prologue writes it), null: copy node instead
    uint32_t flags = 0;
    bool use_wa = false;
    bool fixed = false;               // fixed-cosmology mode (pop_model)
    double* d_fixed_tab = nullptr;    // dVdzdt on the 1024-knot z grid
    bool fixed_tab_set = false;
    cudaStream_t stream = nullptr;
    DataSet evt, sel;
    double ndraw = 1.0;
    // plan
    bool plan_dirty = true;
    Work work{};
    int nrecords = 0, grid = 0, sm_count = 0, lpe = 1, nb_sel = 1;
    int* d_rec_off = nullptr;
    double* d_part = nullptr;
    double* d_slots = nullptr;
    // workspaces
    double *d_theta = nullptr, *d_aux = nullptr, *d_blob = nullptr, *d_partial = nullptr, *d_gather = nullptr,
           *d_out = nullptr;
    unsigned int* d_ticket = nullptr;
    double *h_theta = nullptr, *h_out = nullptr;   // pinned
    unsigned long long* h_done = nullptr;          // pinned: number of completed host-call evaluations (see bump_eval)
    unsigned long long* d_seq = nullptr;           // ... its device-side source, bumped by the epilogue
    unsigned long long host_seq = 0;               // host-call evaluations launched
    cudaGraphExec_t graph_host = nullptr;          // copy in -> 3 kernels -> copy out -> copy of the counter
    int64_t out_len = OUT_HEADER;
    cudaGraphExec_t graph = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // communicator
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    // fused exchange over peer memory
    Mailbox* d_mailbox = nullptr;            // this rank's mailbox (cudaMalloc: IPC-exportable)
    Peers* d_peers = nullptr;                // device copy of the peer table (null: not attached)
    Peers h_peers{};
    unsigned long long* d_epoch = nullptr;   // exchange state: [0] epoch, [1] broken
    double p2p_timeout_s = 10.0;
    void* peer_ptr[P2P_MAX_RANKS] = {};      // mappings opened with cudaIpcOpenMemHandle
    bool slot_counted = false;
    bool slot_pinned = false;                // an evaluation of this context was captured into a caller's graph
    unsigned long long* d_timeline = nullptr;   // bump_debug_timeline
    void *d_arena = nullptr, *d_plan_arena = nullptr;   // the small buffers live in two allocations
};

namespace {

int set_device(const bump_ctx* c) {
    CK(cudaSetDevice(c->device));
    return BUMP_OK;
}

void drop_graphs(bump_ctx* c) {
    if (c->graph) cudaGraphExecDestroy(c->graph), c->graph = nullptr;
    if (c->graph_host) cudaGraphExecDestroy(c->graph_host), c->graph_host = nullptr;
}

void free_plan(bump_ctx* c) {
    drop_graphs(c);
    cudaFree(c->d_plan_arena), c->d_plan_arena = nullptr;
    c->d_rec_off = nullptr, c->d_part = nullptr, c->d_slots = nullptr, c->d_out = nullptr;
    if (c->h_out) cudaFreeHost(c->h_out), c->h_out = nullptr;
}

int upload_set(bump_ctx* c, DataSet& ds, int64_t nrows, int64_t ncols, const double* m1d, const double* q,
               const double* dl, const double* pd) {
    if (nrows < 0 || ncols < 0) return fail(BUMP_E_INVALID, "negative size");
    if (nrows * ncols > 0 && (!m1d || !q || !dl || !pd)) return fail(BUMP_E_INVALID, "null data pointer");
    if (c->fixed && !c->fixed_tab_set)
        return fail(BUMP_E_INVALID, "fixed-cosmology mode: call bump_set_fixed_dvdzdt before uploading data");
    if (int r = set_device(c)) return r;
    ds.reset();
    ds.nrows = nrows;
    ds.ncols = ncols;
    ds.stride = (ncols + GROUP - 1) / GROUP * GROUP;   // whole 64-sample groups: the kernel's loads are unpredicated
    c->plan_dirty = true;
    const int64_t n = nrows * ncols, npad = nrows * ds.stride;
    if (npad == 0) return BUMP_OK;
    if (ds.alloc(sizeof(double) * NCOL * (npad + 2 * GROUP)))   // slack: the kernel looks two blocks ahead
        return fail(BUMP_E_CUDA, "cudaMalloc of the resident columns failed (out of device memory?)");
    double* raw = nullptr;
    CK(cudaMalloc(&raw, sizeof(double) * 4 * n));
    const double* src[4] = {m1d, q, dl, pd};
    for (int k = 0; k < 4; ++k)
        CK(cudaMemcpyAsync(raw + (size_t)k * n, src[k], sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    const int blocks = (int)std::min<int64_t>((npad + 255) / 256, 148 * 16);
    unsigned int* perm = nullptr;
    unsigned long long *keys = nullptr, *keys_out = nullptr;
    unsigned int* idx = nullptr;
    void* tmp = nullptr;
    const bool sort = !(c->flags & BUMP_FLAG_NO_SORT) && n > 64 && n < (int64_t(1) << 32) && nrows < (int64_t(1) << 27);
    if (sort) {
        CK(cudaMalloc(&keys, sizeof(unsigned long long) * n));
        CK(cudaMalloc(&keys_out, sizeof(unsigned long long) * n));
        CK(cudaMalloc(&idx, sizeof(unsigned int) * n));
        CK(cudaMalloc(&perm, sizeof(unsigned int) * n));
        locality_keys_kernel<<<blocks, 256, 0, c->stream>>>(raw, raw + n, raw + 2 * n, ncols, n, keys, idx);
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, idx, perm, n, 0, 64, c->stream));
        CK(cudaMalloc(&tmp, tmp_bytes));
        CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, idx, perm, n, 0, 64, c->stream));
    }
    CK(cudaMemsetAsync(c->d_ticket + 3, 0, sizeof(unsigned int), c->stream));
    prepare_columns_kernel<<<blocks, 256, 0, c->stream>>>(raw, raw + n, raw + 2 * n, raw + 3 * n, perm, nrows, ncols,
                                                          ds.stride, ds.base, c->d_ticket + 3,
                                                          c->fixed ? c->d_fixed_tab : nullptr);
    CK(cudaGetLastError());
    unsigned int bad = 0;
    CK(cudaMemcpyAsync(&bad, c->d_ticket + 3, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(tmp), cudaFree(keys), cudaFree(keys_out), cudaFree(idx), cudaFree(perm);
    CK(cudaFree(raw));
    if (bad) {
        ds.reset();
        return fail(BUMP_E_INVALID, "input arrays must be finite and strictly positive (m1_det, q, d_L, pdraw)");
    }
    return BUMP_OK;
}

int build_plan(bump_ctx* c) {
    if (int r = set_device(c)) return r;
    free_plan(c);
    Work& w = c->work;
    w.nobs = (int32_t)c->evt.nrows;
    w.evt_stride = c->evt.stride;
    w.sel_stride = c->sel.nrows ? c->sel.stride : 0;
    w.g_evt = std::max<int64_t>(1, (w.evt_stride + GROUP - 1) / GROUP);
    w.n_evt_groups = w.nobs * w.g_evt;
    w.n_groups = w.n_evt_groups + (w.sel_stride + GROUP - 1) / GROUP;
    if (w.n_groups >= (int64_t(1) << 31))
        return fail(BUMP_E_INVALID, "shard too large: more than 2^31 groups of 64 samples on one rank");
    // one CTA per SM (persistent); use fewer CTAs only when there are fewer groups than warps
    const int64_t max_warps = (int64_t)c->sm_count * STREAM_WARPS;
    if (const char* e = getenv("BUMP_GPW")) w.gpw = std::max<int64_t>(1, atoll(e));
    else w.gpw = std::max<int64_t>(1, (w.n_groups + max_warps - 1) / max_warps);
    w.nwarps = (int32_t)std::max<int64_t>(1, (w.n_groups + w.gpw - 1) / w.gpw);
    c->grid = (w.nwarps + STREAM_WARPS - 1) / STREAM_WARPS;
    std::vector<int> rec_off(w.nwarps + 1, 0);
    for (int i = 0; i < w.nwarps; ++i) {
        const int64_t g0 = (int64_t)i * w.gpw, g1 = std::min(g0 + w.gpw, w.n_groups);
        const int n = g0 < g1 ? (int)(group_event(w, g1 - 1) - group_event(w, g0) + 1) : 0;
        rec_off[i + 1] = rec_off[i] + n;
    }
    c->nrecords = rec_off[w.nwarps];
    // epilogue: lanes per event = records per event rounded up to a power of two (<= 32)
    const int64_t rpe = (w.g_evt + w.gpw - 1) / w.gpw + 1;
    c->lpe = 1;
    while (c->lpe < 32 && c->lpe < rpe) c->lpe *= 2;
    c->out_len = OUT_HEADER + c->evt.nrows;
    const size_t b_rec = (sizeof(int) * rec_off.size() + 255) / 256 * 256;
    const size_t b_part = (sizeof(double) * PART_STRIDE * std::max(1, c->nrecords) + 255) / 256 * 256;
    const int epb = EPI_THREADS / c->lpe;
    {   // injection blocks of the epilogue: one per 256 warps that own injection groups, at most 8
        const int64_t sel_groups = w.n_groups - w.n_evt_groups;
        const int64_t sel_warps = sel_groups > 0 ? (w.n_groups - 1) / w.gpw - w.n_evt_groups / w.gpw + 1 : 0;
        c->nb_sel = (int)std::min<int64_t>(8, std::max<int64_t>(1, (sel_warps + EPI_THREADS - 1) / EPI_THREADS));
    }
    const size_t b_slots = (sizeof(double) * EPI_SLOT * ((w.nobs + epb - 1) / epb + c->nb_sel) + 255) / 256 * 256;
    const size_t b_out = (sizeof(double) * c->out_len + 255) / 256 * 256;
    CK(cudaMalloc(&c->d_plan_arena, b_rec + b_part + b_slots + b_out));   // one allocation: see bump_ctx_create
    {
        char* base = static_cast<char*>(c->d_plan_arena);
        c->d_out = reinterpret_cast<double*>(base);
        c->d_slots = reinterpret_cast<double*>(base + b_out);
        c->d_rec_off = reinterpret_cast<int*>(base + b_out + b_slots);
        c->d_part = reinterpret_cast<double*>(base + b_out + b_slots + b_rec);
    }
    CK(cudaMallocHost(&c->h_out, sizeof(double) * c->out_len));
    H2D_SYNC(c, c->d_rec_off, rec_off.data(), sizeof(int) * rec_off.size());
    c->plan_dirty = false;
    return BUMP_OK;
}

Columns columns_of(const bump_ctx* c) {
    Columns cols;
    cols.evt_base = c->evt.base;
    cols.sel_base = c->sel.base;
    return cols;
}

EvalConsts consts_of(const bump_ctx* c) {
    EvalConsts ec;
    ec.log_nsamp = log((double)std::max<int64_t>(c->evt.ncols, 1));
    ec.log_ndraw = log(c->ndraw);
    ec.use_wa = c->use_wa ? 1 : 0;
    ec.fixed = c->fixed ? 1 : 0;
    return ec;
}

// stream_kernel<WA, FIXED, SLOT> of a context (its mode and its constant-bank slot)
using StreamKernel = void (*)(const Columns, const Work, const int*, const double*, double*, unsigned long long*);
template <int SLOT>
StreamKernel stream_kernel_of(const bool fixed, const bool wa) {
    return fixed ? stream_kernel<false, true, SLOT> : wa ? stream_kernel<true, false, SLOT> : stream_kernel<false, false, SLOT>;
}
StreamKernel stream_kernel_at(const bool fixed, const bool wa, const int slot) {
    switch (slot) {
        case 0: return stream_kernel_of<0>(fixed, wa);
        case 1: return stream_kernel_of<1>(fixed, wa);
        case 2: return stream_kernel_of<2>(fixed, wa);
        default: return stream_kernel_of<3>(fixed, wa);
    }
}
static_assert(NSLOT == 4, "stream_kernel_at enumerates the slots");
StreamKernel stream_kernel_for(const bump_ctx* c) { return stream_kernel_at(c->fixed, c->use_wa, c->slot); }

// A launch that may begin before the kernel in front of it on the stream has finished (programmatic dependent launch):
// its blocks are scheduled as soon as every block of the predecessor has executed griddepcontrol.launch_dependents and
// an SM has room, and they must execute griddepcontrol.wait before touching anything the predecessor writes (the wait
// returns when the predecessor has completed and its writes are visible).  Inside a captured graph the edge becomes a
// programmatic dependency.  What it buys is the launch latency between two kernels (~1 us each, of a ~45 us
// evaluation at GWTC-3 size).  BUMP_NO_PDL: plain stream order.
template <class... KArgs, class... Args>
cudaError_t launch_dependent(const bool programmatic, void (*kernel)(KArgs...), const dim3 grid, const dim3 block,
                             const size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = programmatic ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// The per-rank part of one evaluation: theta -> partial (+ neff).  3 launches.
int launch_partial(bump_ctx* c, const double* theta_dev, double* partial_dev, double* neff_dev, cudaStream_t s,
                   cudaEvent_t k0 = nullptr, cudaEvent_t k1 = nullptr, double* fused_out = nullptr,
                   unsigned long long* tl = nullptr, const bool count = false) {
    prologue_kernel<<<PRO_BLOCKS, PRO_THREADS, 0, s>>>(theta_dev, c->d_aux, c->d_blob, c->d_cbank, c->d_ticket + 4,
                                                       consts_of(c), tl);
#if !defined(BUMP_SCALARS_FROM_BLOB) && defined(BUMP_CBANK_COPY_NODE)
    CK(cudaMemcpyToSymbolAsync(K_SC4, c->d_blob + OFF_SCAL, sizeof(double) * NSCAL,
                               sizeof(double) * NSCAL * c->slot, cudaMemcpyDeviceToDevice, s));
#endif
    if (k0) cudaEventRecord(k0, s);
    if (c->work.n_groups > 0)
        CK(launch_dependent(PDL_STREAM, stream_kernel_for(c), dim3(c->grid), dim3(STREAM_THREADS),
                            stream_smem_bytes(c->use_wa, c->fixed), s, columns_of(c), c->work, c->d_rec_off, c->d_blob,
                            c->d_part, tl));
    if (k1) cudaEventRecord(k1, s);
    const int epb = EPI_THREADS / c->lpe;
    const int nb_evt = (c->work.nobs + epb - 1) / epb;
    CK(launch_dependent(PDL_EPILOGUE, epilogue_kernel, dim3(nb_evt + c->nb_sel), dim3(EPI_THREADS), 0, s, c->d_part, c->d_rec_off,
                        c->work, (double)c->sel.ncols, c->lpe, c->nb_sel, c->d_blob, neff_dev, c->d_slots,
                        c->d_ticket + 1, partial_dev, fused_out, fused_out ? c->d_peers : nullptr, c->d_epoch, tl,
                        (count && fused_out) ? c->d_seq : nullptr));
    CK(cudaGetLastError());
    return BUMP_OK;
}

int launch_eval(bump_ctx* c, const double* theta_dev, double* out_dev, cudaStream_t s, cudaEvent_t k0 = nullptr,
                cudaEvent_t k1 = nullptr, unsigned long long* tl = nullptr, const bool count = false) {
    // single rank, or peer-memory exchange: the epilogue's last block finalizes in place (3 launches per evaluation)
    if (int r = launch_partial(c, theta_dev, c->d_partial, out_dev + OUT_HEADER, s, k0, k1, c->comm ? nullptr : out_dev, tl,
                               count))
        return r;
    if (c->comm) {
        NCK(g_nccl.AllGather(c->d_partial, c->d_gather, PARTIAL_LEN, NCCL_FLOAT64, c->comm, s));
        finalize_kernel<<<1, 32, 0, s>>>(c->d_gather, c->nranks, out_dev, tl);
        CK(cudaGetLastError());
    }
    return BUMP_OK;
}

// A constant-bank slot (K_SC4[slot]) may be shared by several contexts of a device (more than NSLOT contexts): their
// evaluations are chained through an event so that two of them never have an evaluation in flight at the same time.
// Contexts on different slots run concurrently.  An evaluation that is being CAPTURED into a caller's CUDA graph
// (bump_eval_device on a capturing stream: how an XLA command buffer or torch.cuda.graph calls it) cannot take part
// in that chain - the graph replays later, outside the library's control, and events recorded inside a capture
// cannot be waited on outside it - so capture requires a context that owns its slot alone, and pins the slot: no
// later context is assigned to it.
std::mutex g_dev_mutex;                  // slot assignment
std::mutex g_slot_mutex[64 * NSLOT];     // launch order within a slot
cudaEvent_t g_dev_event[64 * NSLOT] = {};
int g_slot_users[64 * NSLOT] = {};
bool g_slot_pinned[64 * NSLOT] = {};

struct SlotGuard {
    std::unique_lock<std::mutex> lock;
    cudaEvent_t ev = nullptr;
    cudaStream_t s = nullptr;
    // own_stream: the launch goes to the context's own stream (never capturing, and in order with everything else the
    // library launched for this context).  A context that has its slot to itself then needs no event chain at all:
    // three driver calls fewer per evaluation.
    int begin(bump_ctx* c, cudaStream_t stream, const bool own_stream = false) {
        s = stream;
        if (c->device < 0 || c->device >= 64) return BUMP_OK;
        const int k = c->device * NSLOT + c->slot;
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (!own_stream) CK(cudaStreamIsCapturing(stream, &st));
        if (st != cudaStreamCaptureStatusNone) {
            std::lock_guard<std::mutex> lk(g_dev_mutex);
            if (g_slot_users[k] > 1)
                return fail(BUMP_E_INVALID, "stream capture needs a context that owns its constant-bank slot: at most "
                                            "4 contexts per device may be alive when one of them is captured");
            g_slot_pinned[k] = true;
            c->slot_pinned = true;
            return BUMP_OK;   // ordering inside the caller's graph is the caller's stream order
        }
        lock = std::unique_lock<std::mutex>(g_slot_mutex[k]);
        // (a context that makes the slot shared takes this mutex and drains the device first: bump_ctx_create)
        if (own_stream && g_slot_users[k] <= 1) return BUMP_OK;
        if (!g_dev_event[k]) CK(cudaEventCreateWithFlags(&g_dev_event[k], cudaEventDisableTiming));
        ev = g_dev_event[k];
        CK(cudaStreamWaitEvent(s, ev, 0));
        return BUMP_OK;
    }
    ~SlotGuard() {
        if (ev) cudaEventRecord(ev, s);
    }
};

int ensure_ready(bump_ctx* c, cudaStream_t capture_check = nullptr) {
    if (!c) return fail(BUMP_E_INVALID, "null context");
    if (int r = set_device(c)) return r;
    if (c->plan_dirty) {
        if (capture_check) {   // building the plan allocates: not allowed while the caller's stream is capturing
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            CK(cudaStreamIsCapturing(capture_check, &st));
            if (st != cudaStreamCaptureStatusNone)
                return fail(BUMP_E_INVALID, "call bump_plan_info (or one evaluation) after the uploads and before "
                                            "capturing bump_eval_device into a graph");
        }
        if (int r = build_plan(c)) return r;
    }
    return BUMP_OK;
}

int ensure_graph(bump_ctx* c) {
    if (c->graph || (c->flags & BUMP_FLAG_NO_GRAPH)) return BUMP_OK;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int r = launch_eval(c, c->d_theta, c->d_out, c->stream);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (r) {
        if (g) cudaGraphDestroy(g);
        return r;
    }
    CK(e);
    CK(cudaGraphInstantiate(&c->graph, g, 0));
    CK(cudaGraphDestroy(g));
    return BUMP_OK;
}

// The host call as ONE graph: theta from pinned host memory, the three kernels, the result and then the counter of
// completed evaluations to pinned host memory.  bump_eval launches it (one driver call) and polls the counter in
// user space instead of issuing two copies, the launch and a stream synchronisation.
int ensure_graph_host(bump_ctx* c) {
    if (c->graph_host) return BUMP_OK;
    const int nth = c->use_wa ? NTHETA_MAX : NTHETA;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * nth, cudaMemcpyHostToDevice, c->stream);
    int r = (e == cudaSuccess) ? launch_eval(c, c->d_theta, c->d_out, c->stream, nullptr, nullptr, nullptr, true) : BUMP_OK;
    if (e == cudaSuccess && !r)
        e = cudaMemcpyAsync(c->h_out, c->d_out, sizeof(double) * c->out_len, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && !r)
        e = cudaMemcpyAsync(c->h_done, c->d_seq, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
    if (r) {
        if (g) cudaGraphDestroy(g);
        return r;
    }
    CK(e);
    CK(e2);
    CK(cudaGraphInstantiate(&c->graph_host, g, 0));
    CK(cudaGraphDestroy(g));
    return BUMP_OK;
}
bool host_graph_usable(const bump_ctx* c) {
    static const bool on = getenv("BUMP_HOST_GRAPH") != nullptr;
    return on && !(c->flags & BUMP_FLAG_NO_GRAPH) && !c->comm;   // (NCCL exchange: finalize_kernel writes the result)
}

// Wait until the counter in pinned host memory says that host-call evaluation `expect` is complete (spinning in user
// space); the stream is queried now and then so that a faulted kernel ends the wait with its error.
int wait_done(bump_ctx* c, const unsigned long long expect) {
    const volatile unsigned long long* f = c->h_done;
    for (unsigned int it = 1;; ++it) {
        if (*f == expect) break;
        if ((it & 0x1FFFu) == 0u) {
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess) {
                if (*f == expect) break;
                return fail(BUMP_E_CUDA, "evaluation finished without publishing its completion counter");
            }
            if (e != cudaErrorNotReady) return fail(BUMP_E_CUDA, std::string("evaluation failed: ") + cudaGetErrorString(e));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return BUMP_OK;
}

int run_once(bump_ctx* c) {   // d_theta -> d_out on the context stream
    if (!(c->flags & BUMP_FLAG_NO_GRAPH))
        if (int r = ensure_graph(c)) return r;
    SlotGuard chain;
    if (int r = chain.begin(c, c->stream, true)) return r;
    if (c->flags & BUMP_FLAG_NO_GRAPH) return launch_eval(c, c->d_theta, c->d_out, c->stream);
    CK(cudaGraphLaunch(c->graph, c->stream));
    return BUMP_OK;
}

}  // namespace

extern "C" {

const char* bump_version(void) { return "bump_b200 0.1 sm_100a"; }
const char* bump_last_error(void) { return g_err.c_str(); }

int bump_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bump_ctx_create(bump_ctx** out, int device, uint32_t flags) {
    if (!out) return fail(BUMP_E_INVALID, "null out pointer");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return fail(BUMP_E_NOGPU, "no CUDA device visible: bump_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(BUMP_E_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(BUMP_E_NOGPU, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                      "; this library is built for sm_100a only");
    bump_ctx* c = new bump_ctx();
    c->device = device;
    c->flags = flags;
    c->use_wa = (flags & BUMP_FLAG_WA) != 0;
    c->fixed = (flags & BUMP_FLAG_FIXED_COSMO) != 0;
    if (c->fixed && c->use_wa) {
        delete c;
        return fail(BUMP_E_INVALID, "BUMP_FLAG_FIXED_COSMO and BUMP_FLAG_WA are mutually exclusive");
    }
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // ONE allocation for the small per-context buffers: after an O5-size pass has swept the TLBs, the single-warp
        // tails of the prologue and of the epilogue then miss on one page, not on one page per buffer
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t o = off;
            off += (bytes + 255) / 256 * 256;
            return o;
        };
        const size_t o_theta = take(sizeof(double) * NTHETA_MAX), o_aux = take(sizeof(double) * AUX_DOUBLES),
                     o_blob = take(BLOB_BYTES_MAX), o_partial = take(sizeof(double) * PARTIAL_LEN),
                     o_ticket = take(sizeof(unsigned int) * 8), o_fixed = take(sizeof(double) * NZ),
                     o_epoch = take(2 * sizeof(unsigned long long)), o_seq = take(sizeof(unsigned long long)), o_tl = take(sizeof(unsigned long long) * (2 * TL_N + TL_WARP_SLOTS));
        CK(cudaMalloc(&c->d_arena, off));
        CK(cudaMemset(c->d_arena, 0, off));
        CK(cudaDeviceSynchronize());   // (a memset of device memory is asynchronous, and c->stream does not wait for the default stream)
        char* base = static_cast<char*>(c->d_arena);
        c->d_theta = reinterpret_cast<double*>(base + o_theta);
        c->d_aux = reinterpret_cast<double*>(base + o_aux);
        c->d_blob = reinterpret_cast<double*>(base + o_blob);
        c->d_partial = reinterpret_cast<double*>(base + o_partial);
        c->d_ticket = reinterpret_cast<unsigned int*>(base + o_ticket);
        c->d_fixed_tab = reinterpret_cast<double*>(base + o_fixed);
        c->d_epoch = reinterpret_cast<unsigned long long*>(base + o_epoch);
        c->d_timeline = reinterpret_cast<unsigned long long*>(base + o_tl);
        c->d_seq = reinterpret_cast<unsigned long long*>(base + o_seq);
    }
    {   // theta-independent part of the blob: 2^(j/NEXPT), correctly rounded (x87 extended precision on the host)
        std::vector<double> expt(EXPT_DOUBLES);
        for (int j = 0; j < NEXPT; ++j)
            for (int r = 0; r < EXPT_REPL; ++r) expt[j * EXPT_REPL + r] = (double)exp2l((long double)j / NEXPT);
        H2D_SYNC(c, c->d_blob + OFF_EXPT, expt.data(), sizeof(double) * EXPT_DOUBLES);
    }
    // [0] unused, [1] epilogue ticket, [2] unused, [3] bad-input flag, [4] prologue ticket, [5] bad-theta flag
    CK(cudaMallocHost(&c->h_theta, sizeof(double) * NTHETA_MAX));
    CK(cudaMallocHost(&c->h_done, 64));
    *c->h_done = 0ull;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    if (device < 64) {   // the least-used constant-bank slot of this device that no captured graph has pinned
        std::lock_guard<std::mutex> lk(g_dev_mutex);
        int best = -1;
        for (int k = 0; k < NSLOT; ++k)
            if (!g_slot_pinned[device * NSLOT + k] &&
                (best < 0 || g_slot_users[device * NSLOT + k] < g_slot_users[device * NSLOT + best]))
                best = k;
        if (best < 0) {
            bump_ctx_destroy(c);
            return fail(BUMP_E_INVALID, "every constant-bank slot of this device is pinned by a context whose "
                                        "evaluation was captured into a graph: destroy one of them first");
        }
        c->slot = best;
        {   // the slot becomes shared: contexts that had it to themselves launch without the event chain (SlotGuard),
            // so their work in flight is drained once, under the slot's launch mutex, before the count changes
            std::lock_guard<std::mutex> lk2(g_slot_mutex[device * NSLOT + best]);
            if (g_slot_users[device * NSLOT + best] >= 1) cudaDeviceSynchronize();
            ++g_slot_users[device * NSLOT + best];
        }
        c->slot_counted = true;
    }
#if !defined(BUMP_SCALARS_FROM_BLOB) && !defined(BUMP_CBANK_COPY_NODE)
    {
        void* p = nullptr;
        CK(cudaGetSymbolAddress(&p, K_SC4));
        c->d_cbank = static_cast<double*>(p) + (size_t)NSCAL * c->slot;
    }
#endif
    CK(cudaFuncSetAttribute(stream_kernel_for(c), cudaFuncAttributeMaxDynamicSharedMemorySize,
                            stream_smem_bytes(c->use_wa, c->fixed)));
    {
        // every kernel of an evaluation asks for the shared-memory carve-out the streaming kernel needs: an SM does not
        // have to drain and re-partition L1 / shared memory between the four launches (GWTC-3 shape: 56.8 -> 55.3 us per evaluation)
        const int co = cudaSharedmemCarveoutMaxShared;
        CK(cudaFuncSetAttribute(stream_kernel_for(c), cudaFuncAttributePreferredSharedMemoryCarveout, co));
        CK(cudaFuncSetAttribute(prologue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
        CK(cudaFuncSetAttribute(epilogue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
        CK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
    }
    *out = c;
    return BUMP_OK;
}

void bump_ctx_destroy(bump_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->slot_counted) {
        std::lock_guard<std::mutex> lk(g_dev_mutex);
        std::lock_guard<std::mutex> lk2(g_slot_mutex[c->device * NSLOT + c->slot]);
        --g_slot_users[c->device * NSLOT + c->slot];
        if (c->slot_pinned) g_slot_pinned[c->device * NSLOT + c->slot] = false;
    }
    free_plan(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
        if (c->peer_ptr[r]) cudaIpcCloseMemHandle(c->peer_ptr[r]);
    cudaFree(c->d_mailbox);
    cudaFree(c->d_arena);
    cudaFree(c->d_plan_arena);
    cudaFree(c->d_peers);
    c->evt.reset();
    c->sel.reset();
    cudaFree(c->d_gather);
    if (c->h_theta) cudaFreeHost(c->h_theta);
    if (c->h_done) cudaFreeHost(c->h_done);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int bump_upload_events(bump_ctx* c, int64_t nobs, int64_t nsamp, const double* m1s_det, const double* qs,
                       const double* dls, const double* pdraw) {
    if (!c) return fail(BUMP_E_INVALID, "null context");
    return upload_set(c, c->evt, nobs, nsamp, m1s_det, qs, dls, pdraw);
}

int bump_upload_injections(bump_ctx* c, int64_t nsel, const double* m1s_det_sel, const double* qs_sel,
                           const double* dls_sel, const double* pdraw_sel, double ndraw) {
    if (!c) return fail(BUMP_E_INVALID, "null context");
    if (!(ndraw > 0.0)) return fail(BUMP_E_INVALID, "ndraw must be positive");
    c->ndraw = ndraw;
    return upload_set(c, c->sel, nsel > 0 ? 1 : 0, nsel, m1s_det_sel, qs_sel, dls_sel, pdraw_sel);
}

int bump_ctx_clone(bump_ctx* src, bump_ctx** out) {
    if (!src || !out) return fail(BUMP_E_INVALID, "null argument");
    *out = nullptr;
    if (src->d_peers || src->comm) return fail(BUMP_E_INVALID, "clone a context before attaching a communicator");
    bump_ctx* c = nullptr;
    if (int r = bump_ctx_create(&c, src->device, src->flags)) return r;
    if (src->fixed_tab_set) {
        if (cudaMemcpy(c->d_fixed_tab, src->d_fixed_tab, sizeof(double) * NZ, cudaMemcpyDeviceToDevice) != cudaSuccess) {
            bump_ctx_destroy(c);
            return fail(BUMP_E_CUDA, "copy of the fixed-cosmology table failed");
        }
        c->fixed_tab_set = true;
    }
    cudaStreamSynchronize(src->stream);   // uploads of the source are complete
    c->evt = src->evt;                    // shares the resident columns (reference-counted)
    c->sel = src->sel;
    c->ndraw = src->ndraw;
    c->plan_dirty = true;
    *out = c;
    return BUMP_OK;
}

int bump_set_fixed_dvdzdt(bump_ctx* c, const double* dvdzdt, int64_t n) {
    if (!c || !dvdzdt) return fail(BUMP_E_INVALID, "null argument");
    if (!c->fixed) return fail(BUMP_E_INVALID, "context was not created with BUMP_FLAG_FIXED_COSMO");
    if (n != NZ) return fail(BUMP_E_INVALID, "the dVdzdt table must have 1024 entries (zinterp of intensity_models.py:324)");
    if (int r = set_device(c)) return r;
    H2D_SYNC(c, c->d_fixed_tab, dvdzdt, sizeof(double) * NZ);
    c->fixed_tab_set = true;
    return BUMP_OK;
}

int64_t bump_out_len(const bump_ctx* c) { return c ? OUT_HEADER + c->evt.nrows : 0; }

int bump_eval(bump_ctx* c, const double* theta, double* out) {
    if (!theta || !out) return fail(BUMP_E_INVALID, "null theta/out");
    if (int r = ensure_ready(c)) return r;
    const int nth = c->use_wa ? NTHETA_MAX : NTHETA;
    memcpy(c->h_theta, theta, sizeof(double) * nth);
    if (host_graph_usable(c)) {
        if (int r = ensure_graph_host(c)) return r;
        const unsigned long long expect = c->host_seq + 1ull;
        {
            SlotGuard chain;
            if (int r = chain.begin(c, c->stream, true)) return r;
            CK(cudaGraphLaunch(c->graph_host, c->stream));
        }
        c->host_seq = expect;
        if (int r = wait_done(c, expect)) return r;
    } else {
        CK(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * nth, cudaMemcpyHostToDevice, c->stream));
        if (int r = run_once(c)) return r;
        CK(cudaMemcpyAsync(c->h_out, c->d_out, sizeof(double) * c->out_len, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    memcpy(out, c->h_out, sizeof(double) * c->out_len);
    if (c->h_out[OUT_STATUS] != STATUS_OK)
        return fail(BUMP_E_EXCHANGE,
                    c->h_out[OUT_STATUS] == STATUS_EXCHANGE_TIMEOUT
                        ? "multi-rank exchange timed out: a peer did not deliver its partial (this evaluation failed on "
                          "every rank; detach and re-attach the peer mailboxes to continue)"
                        : "multi-rank exchange failed on a peer rank (poisoned mailbox; this evaluation failed on every "
                          "rank; detach and re-attach the peer mailboxes to continue)");
    return BUMP_OK;
}

int bump_eval_device(bump_ctx* c, const double* theta_dev, double* out_dev, void* stream) {
    if (!theta_dev || !out_dev) return fail(BUMP_E_INVALID, "null theta/out");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int r = ensure_ready(c, s)) return r;
    SlotGuard chain;
    if (int r = chain.begin(c, s)) return r;
    return launch_eval(c, theta_dev, out_dev, s);
}

int bump_eval_partial_device(bump_ctx* c, const double* theta_dev, double* partial_dev, double* neff_dev,
                             void* stream) {
    if (!theta_dev || !partial_dev) return fail(BUMP_E_INVALID, "null theta/partial");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int r = ensure_ready(c, s)) return r;
    SlotGuard chain;
    if (int r = chain.begin(c, s)) return r;
    return launch_partial(c, theta_dev, partial_dev, neff_dev ? neff_dev : c->d_out + OUT_HEADER, s);
}

int bump_finalize_device(bump_ctx* c, const double* partials_dev, int nranks, double* out_header_dev, void* stream) {
    if (!c || !partials_dev || !out_header_dev || nranks < 1) return fail(BUMP_E_INVALID, "bad finalize arguments");
    if (int r = set_device(c)) return r;
    finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partials_dev, nranks, out_header_dev, nullptr);
    CK(cudaGetLastError());
    return BUMP_OK;
}

int bump_eval_partial(bump_ctx* c, const double* theta, double* partial, double* neff_local) {
    if (!theta || !partial) return fail(BUMP_E_INVALID, "null theta/partial");
    if (int r = ensure_ready(c)) return r;
    const int nth = c->use_wa ? NTHETA_MAX : NTHETA;
    memcpy(c->h_theta, theta, sizeof(double) * nth);
    CK(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * nth, cudaMemcpyHostToDevice, c->stream));
    {
        SlotGuard chain;
        if (int r = chain.begin(c, c->stream)) return r;
        if (int r = launch_partial(c, c->d_theta, c->d_partial, c->d_out + OUT_HEADER, c->stream)) return r;
    }
    CK(cudaMemcpyAsync(partial, c->d_partial, sizeof(double) * PARTIAL_LEN, cudaMemcpyDeviceToHost, c->stream));
    if (neff_local && c->evt.nrows > 0)
        CK(cudaMemcpyAsync(neff_local, c->d_out + OUT_HEADER, sizeof(double) * c->evt.nrows, cudaMemcpyDeviceToHost,
                           c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return BUMP_OK;
}

int bump_merge_partials(const double* partials, int nranks, double* out_header) {
    if (!partials || !out_header || nranks < 1) return fail(BUMP_E_INVALID, "bad merge arguments");
    finalize_merge(partials, nranks, out_header);
    return BUMP_OK;
}

int bump_nccl_unique_id(void* id128) {
    if (!id128) return fail(BUMP_E_INVALID, "null id");
    if (int r = load_nccl()) return r;
    NCK(g_nccl.GetUniqueId(id128));
    return BUMP_OK;
}

int bump_comm_attach(bump_ctx* c, const void* id128, int nranks, int rank) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(BUMP_E_INVALID, "bad comm arguments");
    if (int r = load_nccl()) return r;
    if (int r = set_device(c)) return r;
    UniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    c->nranks = nranks;
    c->rank = rank;
    cudaFree(c->d_gather);
    CK(cudaMalloc(&c->d_gather, sizeof(double) * PARTIAL_LEN * nranks));
    drop_graphs(c);
    return BUMP_OK;
}

int bump_p2p_export(bump_ctx* c, void* handle64) {
    if (!c || !handle64) return fail(BUMP_E_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (int r = set_device(c)) return r;
    if (!c->d_mailbox) {
        CK(cudaMalloc(&c->d_mailbox, sizeof(Mailbox)));
        CK(cudaMemset(c->d_mailbox, 0, sizeof(Mailbox)));
        CK(cudaMemset(c->d_epoch, 0, 2 * sizeof(unsigned long long)));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->d_mailbox));
    memcpy(handle64, &h, sizeof(h));
    return BUMP_OK;
}

int bump_p2p_attach(bump_ctx* c, const void* handles, int nranks, int rank) {
    if (!c || !handles || nranks < 1 || nranks > P2P_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(BUMP_E_INVALID, "bad p2p arguments (at most 16 ranks)");
    if (!c->d_mailbox) return fail(BUMP_E_INVALID, "call bump_p2p_export first");
    if (c->comm) return fail(BUMP_E_INVALID, "an NCCL communicator is already attached");
    if (int r = set_device(c)) return r;
    Peers p;
    memset(&p, 0, sizeof(p));
    p.nranks = nranks;
    p.rank = rank;
    if (const char* e = getenv("BUMP_P2P_TIMEOUT_S")) c->p2p_timeout_s = std::max(1e-3, atof(e));
    p.timeout_ns = (unsigned long long)(c->p2p_timeout_s * 1e9);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            p.box[r] = c->d_mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, sizeof(h));
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_ptr[r] = ptr;
        p.box[r] = static_cast<Mailbox*>(ptr);
    }
    if (!c->d_peers) CK(cudaMalloc(&c->d_peers, sizeof(Peers)));
    H2D_SYNC(c, c->d_peers, &p, sizeof(p));
    c->h_peers = p;
    c->nranks = nranks;
    c->rank = rank;
    drop_graphs(c);
    return BUMP_OK;
}

int bump_p2p_detach(bump_ctx* c) {
    if (!c) return fail(BUMP_E_INVALID, "null context");
    if (int r = set_device(c)) return r;
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < P2P_MAX_RANKS; ++r)
        if (c->peer_ptr[r]) cudaIpcCloseMemHandle(c->peer_ptr[r]), c->peer_ptr[r] = nullptr;
    cudaFree(c->d_peers), c->d_peers = nullptr;
    // a fresh mailbox and exchange state (epoch 0, not broken): after a failed exchange every rank detaches, the ranks
    // synchronise on the host, and then export / attach again
    if (c->d_mailbox) CK(cudaMemset(c->d_mailbox, 0, sizeof(Mailbox)));
    if (c->d_epoch) CK(cudaMemset(c->d_epoch, 0, 2 * sizeof(unsigned long long)));
    CK(cudaDeviceSynchronize());
    c->nranks = 1;
    c->rank = 0;
    drop_graphs(c);
    return BUMP_OK;
}

int bump_p2p_set_timeout(bump_ctx* c, double seconds) {
    if (!c || !(seconds > 0.0)) return fail(BUMP_E_INVALID, "bad timeout");
    if (int r = set_device(c)) return r;
    c->p2p_timeout_s = seconds;
    if (c->d_peers) {
        CK(cudaStreamSynchronize(c->stream));
        c->h_peers.timeout_ns = (unsigned long long)(seconds * 1e9);
        H2D_SYNC(c, c->d_peers, &c->h_peers, sizeof(Peers));
    }
    return BUMP_OK;
}

int bump_debug_tables(bump_ctx* c, int which, double* out, int64_t out_len) {
    if (!c || !out) return fail(BUMP_E_INVALID, "null argument");
    if (int r = set_device(c)) return r;
    const double* src = nullptr;
    int64_t n = 0;
    switch (which) {
        case 0: src = c->d_aux + AUX_ZG, n = 4 * NZ; break;
        case 1: src = c->d_aux + AUX_TAN, n = 9 * NZ; break;
        case 2: src = c->d_aux + AUX_G, n = 6 * NM; break;
        case 3: src = c->d_blob + OFF_SCAL, n = NSCAL; break;
        default: return fail(BUMP_E_INVALID, "unknown table id");
    }
    if (out_len < n) return fail(BUMP_E_INVALID, "output buffer too small");
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return BUMP_OK;
}

int bump_debug_math(bump_ctx* c, int which, const double* x, int64_t n, double* y) {
    if (!c || !x || !y || n < 1 || which < 0 || which > 4) return fail(BUMP_E_INVALID, "bad math-probe arguments");
    if (int r = set_device(c)) return r;
    double *dx = nullptr, *dy = nullptr;
    CK(cudaMalloc(&dx, sizeof(double) * n));
    CK(cudaMalloc(&dy, sizeof(double) * n));
    H2D_SYNC(c, dx, x, sizeof(double) * n);
    const int smem = (OFF_EXPT + EXPT_DOUBLES) * 8;
    CK(cudaFuncSetAttribute(math_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    math_probe_kernel<<<64, 256, smem, c->stream>>>(which, dx, n, c->d_blob, dy);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx);
    cudaFree(dy);
    return BUMP_OK;
}

int bump_time_evals(bump_ctx* c, const double* theta, int iters, float* total_ms, float* stream_ms) {
    if (!theta || iters < 1 || !total_ms) return fail(BUMP_E_INVALID, "bad timing arguments");
    if (int r = ensure_ready(c)) return r;
    const int nth = c->use_wa ? NTHETA_MAX : NTHETA;
    memcpy(c->h_theta, theta, sizeof(double) * nth);
    CK(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * nth, cudaMemcpyHostToDevice, c->stream));
    if (int r = ensure_graph(c)) return r;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < iters; ++i)
        if (int r = run_once(c)) return r;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
    if (stream_ms) {   // the streaming kernel alone: events around each direct launch on the same stream
        float acc = 0.f;
        for (int i = 0; i < iters; ++i) {
            SlotGuard chain;
            if (int r = chain.begin(c, c->stream)) return r;
            if (int r = launch_eval(c, c->d_theta, c->d_out, c->stream, c->ev0, c->ev1)) return r;
            CK(cudaEventSynchronize(c->ev1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            acc += ms;
        }
        CK(cudaStreamSynchronize(c->stream));
        *stream_ms = acc;
    }
    return BUMP_OK;
}

int bump_ctx_flags(const bump_ctx* c) { return c ? (int)c->flags : -1; }

int bump_set_error(int code, const char* msg) { return fail(code, msg ? msg : ""); }   // for bump_nuts.cpp

int bump_launches_per_eval(const bump_ctx* c) { return c ? (c->comm ? 4 : 3) : 0; }

int bump_debug_warp_times(bump_ctx* c, double* out_us, int64_t nwarps) {
    if (!c || !out_us || nwarps < 1 || nwarps > TL_WARP_SLOTS) return fail(BUMP_E_INVALID, "bad warp-times arguments");
    if (int r = set_device(c)) return r;
    std::vector<unsigned long long> t(2 * TL_N + nwarps);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(t.data(), c->d_timeline, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
    const unsigned long long t0 = t[2 * TL_STREAM];
    for (int64_t w = 0; w < nwarps; ++w) out_us[w] = t[2 * TL_N + w] ? (double)(t[2 * TL_N + w] - t0) * 1e-3 : -1.0;
    return BUMP_OK;
}

int bump_debug_timeline(bump_ctx* c, const double* theta, double* out_us, int64_t out_len) {
    if (!theta || !out_us || out_len < 2 * TL_N) return fail(BUMP_E_INVALID, "bad timeline arguments");
    if (int r = ensure_ready(c)) return r;
    unsigned long long init[2 * TL_N], got[2 * TL_N];
    for (int k = 0; k < TL_N; ++k) init[2 * k] = ~0ull, init[2 * k + 1] = 0ull;
    const int nth = c->use_wa ? NTHETA_MAX : NTHETA;
    memcpy(c->h_theta, theta, sizeof(double) * nth);
    CK(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * nth, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_timeline, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    if (c->flags & BUMP_FLAG_NO_GRAPH) {
        SlotGuard chain;
        if (int r = chain.begin(c, c->stream)) return r;
        if (int r = launch_eval(c, c->d_theta, c->d_out, c->stream, nullptr, nullptr, c->d_timeline)) return r;
    } else {   // as one graph, like a normal evaluation: the gaps between the kernels are those of a graph replay
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int r = launch_eval(c, c->d_theta, c->d_out, c->stream, nullptr, nullptr, c->d_timeline);
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (r) {
            if (g) cudaGraphDestroy(g);
            return r;
        }
        CK(e);
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphDestroy(g));
        {
            SlotGuard chain;
            if (int r2 = chain.begin(c, c->stream)) return r2;
            CK(cudaGraphLaunch(ge, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGraphExecDestroy(ge));
    }
    CK(cudaMemcpyAsync(got, c->d_timeline, sizeof(got), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long t0 = ~0ull;
    for (int k = 0; k < TL_N; ++k)
        if (got[2 * k + 1] != 0ull) t0 = std::min(t0, got[2 * k]);
    for (int k = 0; k < TL_N; ++k) {
        const bool ran = got[2 * k + 1] != 0ull;
        out_us[2 * k] = ran ? (double)(got[2 * k] - t0) * 1e-3 : -1.0;
        out_us[2 * k + 1] = ran ? (double)(got[2 * k + 1] - t0) * 1e-3 : -1.0;
    }
    return BUMP_OK;
}

int bump_plan_info(bump_ctx* c, int64_t* info8) {
    if (!info8) return fail(BUMP_E_INVALID, "null info");
    if (int r = ensure_ready(c)) return r;
    info8[0] = c->work.n_groups;
    info8[1] = c->work.gpw;
    info8[2] = c->nrecords;
    info8[3] = c->grid;
    info8[4] = STREAM_THREADS;
    info8[5] = stream_smem_bytes(c->use_wa, c->fixed);
    info8[6] = c->evt.nrows * c->evt.stride + c->sel.nrows * c->sel.stride;
    info8[7] = c->sm_count;
    return BUMP_OK;
}

}  // extern "C"

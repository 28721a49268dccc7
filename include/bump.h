/*
 * bump.h — C ABI of the B200-native BumpCosmology hyperlikelihood (libbump_b200.so).
 *
 * The reference has no FFI: its hot path is Python inlined in a numpyro model.  Each entry point below
 * therefore names the reference *Python* interface it replaces (file:line under /root/reference/src/scripts).
 * Plain pointers and sizes only; no torch / CUDA types.  All functions return 0 on success or a BUMP_E_*
 * code, never throw, and leave a thread-local message readable through bump_last_error().
 *
 * theta layout (the derived parameters the reference's density objects receive, intensity_models.py:368-376):
 *   [0] h  [1] Om  [2] w  [3] a  [4] b  [5] c  [6] mpisn  [7] mbhmax  [8] sigma  [9] fpl  [10] beta
 *   [11] lam  [12] kappa  [13] zp          (BUMP_NTHETA = 14)
 * and, only in the w0-wa extension mode, [14] wa (BASELINE.json config 5; no reference counterpart).
 */
#ifndef BUMP_H_
#define BUMP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BUMP_NTHETA 14
#define BUMP_NTHETA_MAX 15 /* with wa */

/* Layout of the flat result vector written by bump_eval (doubles). */
#define BUMP_OUT_LOGLIKE 0      /* sum_i [logsumexp_j w_ij - log nsamp]      'loglike' factor, intensity_models.py:382-383 */
#define BUMP_OUT_LOG_MU_SEL 1   /* logsumexp_k w_k - log Ndraw               intensity_models.py:389 ('selfactor' = -nobs*this, :390) */
#define BUMP_OUT_LOG_MU2 2      /* logsumexp_k 2 w_k - 2 log Ndraw           intensity_models.py:392 */
#define BUMP_OUT_NEFF_SEL 3     /* exp(2 log_mu_sel - log_s2)                intensity_models.py:393-394 */
#define BUMP_OUT_DLOGLIKE 4     /* d loglike / d theta[0..14]   (15 slots; slot 14 = wa, 0 unless wa mode) */
#define BUMP_OUT_DLOG_MU 19     /* d log_mu_sel / d theta[0..14] */
#define BUMP_OUT_NVALID_EVT 34  /* number of event samples with finite weight (diagnostic) */
#define BUMP_OUT_NVALID_SEL 35  /* number of injections with finite weight (diagnostic) */
#define BUMP_OUT_NOBS 36        /* total number of events over all ranks */
#define BUMP_OUT_NSEL 37        /* total number of found injections over all ranks */
#define BUMP_OUT_STATUS 38      /* 0 = ok; 1 / 2 = the multi-rank peer-memory exchange timed out / was poisoned by a
                                 * failing peer (outputs are NaN then, on EVERY rank; bump_eval returns BUMP_E_EXCHANGE) */
#define BUMP_OUT_HEADER 40      /* neff[nobs_local] follows (intensity_models.py:401) */

/* Per-rank partial for the multi-GPU exchange (doubles); see DESIGN.md "multi-GPU". */
#define BUMP_PARTIAL_LEN 128

#define BUMP_OK 0
#define BUMP_E_INVALID 1   /* bad argument / call order */
#define BUMP_E_CUDA 2      /* CUDA runtime error (message has the cudaError string) */
#define BUMP_E_NOGPU 3     /* no CUDA device: there is NO CPU fallback */
#define BUMP_E_NCCL 4      /* NCCL missing or failed */
#define BUMP_E_EXCHANGE 5  /* the peer-memory exchange of a multi-rank evaluation failed (timeout / poisoned): the
                            * evaluation did not happen on ANY rank; detach, synchronise the ranks, attach again */

/* Evaluation flags (bump_ctx_create). */
#define BUMP_FLAG_WA 1u        /* w0-wa (CPL) dark energy: theta has 15 entries */
#define BUMP_FLAG_NO_GRAPH 2u  /* launch kernels directly instead of replaying a CUDA graph */
#define BUMP_FLAG_NO_SORT 4u   /* keep the caller's sample order inside each event (default: locality sort at upload) */
#define BUMP_FLAG_FIXED_COSMO 8u /* fixed cosmology = the reference's pop_model (intensity_models.py:313-355): the upload
                                  * functions then take SOURCE-frame (m1s, qs, zs, pdraw) and theta[0..2] are ignored */

typedef struct bump_ctx bump_ctx;

/* Library / build information: returns e.g. "bump_b200 0.1 sm_100a". */
const char* bump_version(void);
const char* bump_last_error(void);
int bump_device_count(void);

/* One context = one GPU = one shard of the catalog.  Replaces the implicit XLA executable that numpyro builds
 * around pop_cosmo_model (intensity_models.py:357; run_cosmo_fit.py:45-49). */
int bump_ctx_create(bump_ctx** ctx, int device, uint32_t flags);
void bump_ctx_destroy(bump_ctx* ctx);

/* A second context on the SAME resident catalog: the columns uploaded to `src` are shared (read-only, reference
 * counted - one copy in HBM, freed with the last context), everything theta-dependent is the clone's own, so `src` and
 * its clones evaluate different theta concurrently (one per NUTS chain: run_cosmo_fit.py:46 runs 4 chains on one data
 * set).  Uploading to a clone detaches it from the shared columns.  Clone before attaching a communicator. */
int bump_ctx_clone(bump_ctx* src, bump_ctx** clone);

/* Upload this rank's events: four row-major [nobs, nsamp] float64 HOST arrays exactly as run_cosmo_fit.py:32-43
 * builds them (m1s_det, qs, dls, pdraw).  theta-independent logs are precomputed on the device here
 * (the reference recomputes log(pdraw) every trace, intensity_models.py:365).  nobs may be 0. */
int bump_upload_events(bump_ctx* ctx, int64_t nobs, int64_t nsamp, const double* m1s_det, const double* qs,
                       const double* dls, const double* pdraw);

/* Upload this rank's found injections: four [nsel] float64 HOST arrays and the TOTAL number of draws
 * (m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw of intensity_models.py:357; run_cosmo_fit.py:49). */
int bump_upload_injections(bump_ctx* ctx, int64_t nsel, const double* m1s_det_sel, const double* qs_sel,
                           const double* dls_sel, const double* pdraw_sel, double ndraw);

/* Fixed-cosmology mode only, before the uploads: the theta-independent table dVdzdt_interp (1024 doubles) on
 * zinterp = expm1(linspace(log1p(0), log1p(100), 1024)) that pop_model builds from astropy's Planck18
 * (intensity_models.py:323-325); log interp(z, zinterp, dVdzdt_interp) (:332) is folded into the samples at upload. */
int bump_set_fixed_dvdzdt(bump_ctx* ctx, const double* dvdzdt_interp, int64_t n);

/* Number of doubles bump_eval writes: BUMP_OUT_HEADER + nobs_local. */
int64_t bump_out_len(const bump_ctx* ctx);

/* One evaluation of the hot path (intensity_models.py:374-394,401 + its reverse pass): theta (HOST, 14 or 15
 * doubles) -> out (HOST, bump_out_len doubles).  Synchronous: copies theta in, replays the kernel graph,
 * copies the result out (environment BUMP_HOST_GRAPH=1: all of that as one graph launch, the calling thread polling
 * a completion counter in pinned memory instead of synchronising the stream; measured equal, +2 % for four chains).  Non-finite or out-of-support theta yields NaN/-inf outputs, not an error
 * (NUTS relies on that).  If a communicator is attached, the result is the merged all-rank value, identical
 * bit for bit on every rank; neff[] stays local to the rank's events. */
int bump_eval(bump_ctx* ctx, const double* theta, double* out);

/* Same, fully asynchronous on a caller stream with DEVICE pointers (what an XLA FFI handler calls):
 * no allocation, no host synchronisation.  stream is a cudaStream_t passed as void*.  The stream may be CAPTURING
 * (an XLA command buffer, torch.cuda.graph): the evaluation then becomes 3 kernel nodes of the caller's graph (the
 * epilogue as a programmatic dependent of the stream kernel).  Capture needs (a) the plan built beforehand (bump_plan_info or one evaluation after the uploads)
 * and (b) a context that owns its constant-bank slot alone, i.e. at most 4 contexts alive on the device; the slot
 * stays reserved for the captured context until it is destroyed.  A context is not re-entrant: do not run two of its
 * evaluations (captured or not) concurrently.  out_dev[BUMP_OUT_STATUS] reports a failed multi-rank exchange. */
int bump_eval_device(bump_ctx* ctx, const double* theta_dev, double* out_dev, void* stream);

/* Multi-GPU, host-driven exchange (torch.distributed or any allgather): produce this rank's partial
 * (BUMP_PARTIAL_LEN doubles, HOST), then merge nranks partials in rank order into the final header
 * (BUMP_OUT_HEADER doubles, HOST).  The merge is a pure function: every rank computes identical bits. */
int bump_eval_partial(bump_ctx* ctx, const double* theta, double* partial, double* neff_local);
int bump_merge_partials(const double* partials, int nranks, double* out_header);

/* Multi-GPU, caller-driven exchange ON THE DEVICE (torch.distributed all_gather_into_tensor between the two
 * calls, everything on one caller stream, no host synchronisation): this rank's partial (BUMP_PARTIAL_LEN
 * doubles, DEVICE) and neff (nobs_local doubles, DEVICE; may be NULL), then the rank-ordered merge of the
 * gathered [nranks][BUMP_PARTIAL_LEN] partials into the result header (BUMP_OUT_HEADER doubles, DEVICE). */
int bump_eval_partial_device(bump_ctx* ctx, const double* theta_dev, double* partial_dev, double* neff_dev,
                             void* stream);
int bump_finalize_device(bump_ctx* ctx, const double* partials_dev, int nranks, double* out_header_dev,
                         void* stream);

/* Multi-GPU, in-library exchange: one ncclAllGather of the partial + the same merge on the device, inside the
 * evaluation graph.  id is the 128-byte ncclUniqueId produced by bump_nccl_unique_id on rank 0 and broadcast
 * by the host (e.g. torch.distributed). */
int bump_nccl_unique_id(void* id128);
int bump_comm_attach(bump_ctx* ctx, const void* id128, int nranks, int rank);

/* Multi-GPU, fused exchange over peer memory (NVLink/NVSwitch, ranks = processes of ONE node, one GPU each): the last
 * block of the epilogue kernel stores this rank's partial straight into every peer's mailbox, publishes an epoch
 * flag, waits for the peers' flags and merges — no NCCL call and no extra launch on the evaluation path.
 * bump_p2p_export writes this rank's 64-byte cudaIpcMemHandle_t; the host all-gathers the handles (any transport)
 * and passes the [nranks][64] array to bump_p2p_attach.  At most 16 ranks.  Every rank must then call bump_eval /
 * bump_eval_device the same number of times (it is a collective).  A peer that does not arrive within the timeout
 * (default 10 s; bump_p2p_set_timeout or the environment variable BUMP_P2P_TIMEOUT_S) makes the evaluation FAIL ON
 * EVERY RANK, the late one included: the waiting ranks give up, poison their flags in all peer mailboxes and mark
 * their exchange broken; bump_eval returns BUMP_E_EXCHANGE (bump_eval_device: out[BUMP_OUT_STATUS] != 0, NaN outputs)
 * from then on until every rank has called bump_p2p_detach, the ranks have synchronised on the host, and the mailboxes
 * are exported and attached again.  No rank ever sees a result that another rank did not also see. */
int bump_p2p_export(bump_ctx* ctx, void* handle64);
int bump_p2p_attach(bump_ctx* ctx, const void* handles, int nranks, int rank);
int bump_p2p_detach(bump_ctx* ctx);   /* back to a single-rank context (e.g. to fall back to bump_comm_attach) */
int bump_p2p_set_timeout(bump_ctx* ctx, double seconds);

/* Introspection for unit-level parity tests of the prologue kernels (F1-F3 of SURVEY.md section 2.2):
 * copies the theta-dependent tables of the last evaluation to the host.
 *   which = 0: cosmology knots  [4][1024]: zinterp, dlinterp, ddlinterp, dvcinterp      (intensity_models.py:230-235)
 *   which = 1: cosmology tangents [3 tables][2 or 3 params][1024]: d{dl,ddl,dvc}/d{Om,w[,wa]}
 *   which = 2: PISN table [1 + 5][256]: log_dN_grid and its tangents d/d{a,b,mpisn,mbhmax,sigma} (:96-108)
 *   which = 3: scalars [32]: log_pl_norm, log_norm, rate_log_norm, then their tangents */
int bump_debug_tables(bump_ctx* ctx, int which, double* out, int64_t out_len);

/* Accuracy probe of the streaming kernel's own scalar math, y[i] = f(x[i]) on the device (HOST arrays):
 *   which = 0: exp, one-constant reduction (mass function, rate)   1: exp, two-constant reduction (rescaling)
 *           2: reciprocal of a positive normal double               3: log1p on [0, 0.00453]   4: 1/(1+x) on the same */
int bump_debug_math(bump_ctx* ctx, int which, const double* x, int64_t n, double* y);

/* Timing helper for bench.py: run `iters` evaluations back to back on the context's stream with theta
 * already on the device, bracketed by CUDA events ON THAT STREAM; returns total milliseconds and, if
 * stream_ms is non-NULL, the milliseconds spent in the streaming kernel alone (events around each launch,
 * graph replay disabled for that measurement). */
int bump_time_evals(bump_ctx* ctx, const double* theta, int iters, float* total_ms, float* stream_ms);

/* Per-kernel timeline of ONE evaluation launched directly (no graph): out_us[2k], out_us[2k+1] = start of the first
 * block / end of the last block of phase k (0 prologue, 1 streaming, 2 epilogue, 3 finalize, then sub-phases:
 * 4 prologue PISN rows, 5 prologue cosmology blocks, 6 prologue last block, 7 streaming kernel past the table staging,
 * 8 epilogue per-event phase, 9 epilogue last block, 10 first / last warp of the streaming kernel to finish;
 * -1 if it did not run)
 * in microseconds on the GPU's global timer, relative to the start of the first kernel; launched as one CUDA graph
 * like a normal evaluation (directly with BUMP_FLAG_NO_GRAPH).  out_len >= 22. */
int bump_debug_timeline(bump_ctx* ctx, const double* theta, double* out_us, int64_t out_len);
/* After bump_debug_timeline: when each of the first nwarps (<= 4096) warps of the streaming kernel finished, in
 * microseconds from the kernel's start (-1: the warp had no work).  Shows the load balance of the static plan. */
int bump_debug_warp_times(bump_ctx* ctx, double* out_us, int64_t nwarps);

/* Number of kernel launches one bump_eval performs (for bench.py's gpu_launches). */
int bump_launches_per_eval(const bump_ctx* ctx);

/* Execution plan of the streaming kernel: info8 = {64-sample groups, groups per warp, records, grid (CTAs), threads
 * per CTA, dynamic shared memory bytes, padded samples resident on this rank, SM count}. */
int bump_plan_info(bump_ctx* ctx, int64_t* info8);

/* ---- Host NUTS driver.  Replaces what numpyro runs for the reference:
 * `NUTS(pop_cosmo_model, dense_mass=True)`, `MCMC(num_warmup=1000, num_samples=1000, num_chains=4)` with seed
 * 1652819403 (run_cosmo_fit.py:17-19,45-49).  One chain per call; run the chains from concurrent host threads, one
 * context each (contexts of one device evaluate concurrently).  Priors / transforms of the 15 sample sites:
 * intensity_models.py:281-311,398.  Outputs (row-major, caller-allocated):
 *   out_u, out_x [num_samples][15]   unconstrained / constrained draws, site order h Om w a b c mpisn dmbhmax sigma
 *                                    beta log_fpl lam dkappa zp R_unit
 *   out_stats    [num_samples][BUMP_NUTS_NSTAT]   accept prob, tree depth, leapfrog steps, diverging, potential
 *   out_det      [num_samples][BUMP_NUTS_NDET]    loglike, selfactor, neff_sel, R, mbhmax, fpl, kappa, min neff
 *   out_info     [8]   step size, leapfrog steps (total), warm-up seconds, sampling seconds, model evaluations
 *   out_minv     [15][15]  adapted inverse mass matrix (may be NULL; so may out_x / out_det / init_u) */
#define BUMP_NUTS_NSTAT 5
#define BUMP_NUTS_NDET 8
int bump_nuts_chain(bump_ctx* ctx, int num_warmup, int num_samples, uint64_t seed, int dense_mass, double target_accept,
                    int max_tree_depth, const double* init_u, double* out_u, double* out_x, double* out_stats,
                    double* out_det, double* out_info, double* out_minv);
/* What the driver integrates, exposed for parity tests against the reference-minted potential golden
 * (tests/golden/potential_small.npz): the potential energy of pop_cosmo_model in unconstrained space,
 *   U(u) = -[sum_i log prior_i(x_i(u_i)) + sum_i log|dx_i/du_i| + loglike + selfactor]   (numpyro's potential_energy)
 * and dU/du for the 15 sites; rec[BUMP_NUTS_NDET] (may be NULL) = the deterministics of that evaluation.
 * bump_nuts_prior_terms is the prior + Jacobian part alone (no GPU): x[15] = constrained values, *prior_u and
 * prior_grad[15] = that part of U and of dU/du. */
int bump_nuts_potential(bump_ctx* ctx, const double* u, double* U, double* grad, double* rec);
int bump_nuts_prior_terms(const double* u, double* x, double* prior_u, double* prior_grad);
/* The same sampler on an arbitrary potential U(u) with gradient (dim <= 32): the CPU-testable core. */
typedef double (*bump_potential_cb)(void* user, const double* u, double* grad);
int bump_nuts_chain_cb(bump_potential_cb f, void* user, int dim, int num_warmup, int num_samples, uint64_t seed,
                       int dense_mass, double target_accept, int max_tree_depth, const double* init_u, double* out_u,
                       double* out_stats, double* out_info, double* out_minv);
/* The flags the context was created with (-1 for a null context). */
int bump_ctx_flags(const bump_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BUMP_H_ */

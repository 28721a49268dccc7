// Shared constants and memory layouts (host + device).
#pragma once
#include <stdint.h>

namespace bump {

// ---- model constants: file:line in /root/reference/src/scripts/intensity_models.py
constexpr double MBH_MIN = 5.0;        // :13
constexpr double MTR = 20.0;           // :41
constexpr double TURNON_WIDTH = 0.05;  // :45
constexpr int NM = 256;                // :92  n_m
constexpr double MIN_BH_MASS = 3.0;    // :97
constexpr double MIN_CO_MASS = 1.0;    // :98
constexpr double MREF = 30.0;          // :129
constexpr double QREF = 1.0;           // :192
constexpr double ZMAX = 100.0;         // :220
constexpr int NZ = 1024;               // :221 ninterp
constexpr double C_H100_GPC = 2.99792; // :239

constexpr double LOG_ZMAX1 = 4.6151205168412597;            // log(101)
constexpr double ZSTEP = LOG_ZMAX1 / (NZ - 1);              // uniform step of the z grid in log(1+z)
constexpr double LOG_MREF_PAIR = 4.0943445622221004;        // log(mref*(1+qref)) = log 60
constexpr double LOG_MREF = 3.4011973816621555;             // log 30
constexpr double LN2 = 0.69314718055994530942;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;     // log(sqrt(2 pi))
constexpr double FOUR_PI = 12.566370614359172954;

constexpr double MASS_BEYOND_LOG = -5.0e4;   // log dN of the mass table's beyond-the-grid record (exp saturates: ~0)

constexpr int NTHETA = 14;
constexpr int NTHETA_MAX = 15;
enum ThetaIdx { T_H = 0, T_OM, T_W, T_A, T_B, T_C, T_MPISN, T_MBHMAX, T_SIGMA, T_FPL, T_BETA, T_LAM, T_KAPPA,
                T_ZP, T_WA };

// ---- theta-dependent table blob: built by the prologue kernel in global memory, bulk-copied (TMA) into the
// shared memory of every CTA of the streaming kernel.  Offsets in doubles.
constexpr int NSCAL = 64;
constexpr int NCPAIR = 4;    // cosmology per-bin records {f_b, slope-like}
constexpr int NCTAN = 9;     // cosmology tangent tables, one double per knot
constexpr int NMREC = 6;     // mass per-bin records {g_b, g_{b+1}-g_b}
// bucket table for the d_L search: key = top bits of the double (11 exponent bits + SRCH_MBITS mantissa bits)
constexpr int SRCH_MBITS = 8;
constexpr int SRCH_EXP_LO = -8 + 1023;    // biased exponent of 2^-8 Gpc
constexpr int SRCH_OCTAVES = 21;          // 2^-8 .. 2^13 Gpc
constexpr int SRCH_N = SRCH_OCTAVES << SRCH_MBITS;   // 5376 uint16 entries
constexpr int SRCH_DOUBLES = SRCH_N / 4;

// exp table (bump_math.cuh fexp): 2^(j/NEXPT), theta-independent, written once per context.  Each entry is stored
// EXPT_REPL times in a row and lane l reads copy (l mod EXPT_REPL): the 32 lanes of a lookup at random j then spread
// over the banks by construction (EXPT_REPL = 16: each of the 16 64-bit bank pairs serves exactly two lanes = the
// minimum of 2 wavefronts; one copy of a 2048-entry table: ~6).  A shorter table needs a longer polynomial.
// Measured on B200 (quarter-size O5, ns per sample): 2048 x 1: 0.0200, 1024 x 4: 0.0195, 256 x 16: 0.0183 although the
// last needs one more FP64 instruction per exp - the kernel is bound by shared-memory wavefronts, not by issue slots.
#ifndef BUMP_EXPT_LOG2
#define BUMP_EXPT_LOG2 8
#endif
#ifndef BUMP_EXPT_REPL_LOG2
#define BUMP_EXPT_REPL_LOG2 4
#endif
constexpr int NEXPT = 1 << BUMP_EXPT_LOG2;
constexpr int EXPT_REPL = 1 << BUMP_EXPT_REPL_LOG2;
constexpr int EXPT_DOUBLES = NEXPT * EXPT_REPL;
static_assert(EXPT_REPL <= 16, "16 bank pairs");
constexpr int OFF_SCAL = 0;
constexpr int OFF_EXPT = OFF_SCAL + NSCAL;              // double  expt[NEXPT][EXPT_REPL]
constexpr int OFF_COS = OFF_EXPT + EXPT_DOUBLES;        // double2 cos[NCPAIR][NZ]
constexpr int OFF_SRCH = OFF_COS + NCPAIR * NZ * 2;     // uint16  srch[SRCH_N]
constexpr int OFF_MASS = OFF_SRCH + SRCH_DOUBLES;       // double2 mass[NMREC][NM]
// The cosmology tangent tables come last because their format depends on the mode (227 KB of shared memory do not
// hold nine pair tables):
//   default (h, Om, w):  double2 ctan2[6][NZ]  per-bin pairs {t_b, t_{b+1} - t_b}: one LDS.128 and no subtraction per lerp
//   w0-wa             :  double  ctan[9][NZ]   knot values (two LDS.64 and one DADD per lerp)
//   fixed cosmology   :  none (pop_model has no cosmological gradient)
constexpr int OFF_CTAN = OFF_MASS + NMREC * NM * 2;
constexpr int NCTAN_PAIR = 6;
__host__ __device__ constexpr int blob_doubles(const bool wa, const bool fixed) {
    return OFF_CTAN + (fixed ? 0 : wa ? NCTAN * NZ : NCTAN_PAIR * NZ * 2);
}
constexpr int BLOB_DOUBLES_MAX = OFF_CTAN + NCTAN_PAIR * NZ * 2;
constexpr int BLOB_BYTES_MAX = BLOB_DOUBLES_MAX * 8;
static_assert(blob_doubles(true, false) <= BLOB_DOUBLES_MAX, "the pair layout is the largest blob");
static_assert(BLOB_BYTES_MAX <= 227 * 1024, "the blob must fit the 227 KB of shared memory per CTA (the streaming "
                                            "kernel's mbarrier sits in the unused scalar block at its start)");
static_assert(OFF_CTAN % 2 == 0 && (blob_doubles(false, false) % 2) == 0 && (blob_doubles(true, false) % 2) == 0,
              "bulk copies need 16-byte multiples");
static_assert(SRCH_N % 4 == 0, "search table must fill whole doubles");
static_assert(OFF_EXPT % 16 == 0, "the exp table must sit on a 128-byte boundary");
static_assert((NEXPT & (NEXPT - 1)) == 0, "the exp table size is a power of two");

// cosmology pair records, bin b = [knot b, knot b+1]
enum CosRec { CR_DL = 0,   // {dl_b, 1/(dl_{b+1}-dl_b)}
              CR_DVC,      // {dvc_b, dvc_{b+1}-dvc_b}
              CR_DDL,      // {ddl_b, ddl_{b+1}-ddl_b}
              CR_Z };      // {1/(1+z_b), log(1+z_b)}     theta-independent
// cosmology tangent tables (knot values)
enum CosTan { CT_DL_OM = 0, CT_DVC_OM, CT_DDL_OM, CT_DL_W, CT_DVC_W, CT_DDL_W, CT_DL_WA, CT_DVC_WA, CT_DDL_WA };
// mass records, bin b of the mbh grid
enum MassRec { MR_G = 0, MR_GA, MR_GB, MR_GMPISN, MR_GMBHMAX, MR_GSIGMA };

// scalars
enum Scal {
    S_H = 0, S_INV_H, S_C, S_M, S_LOG_M, S_INV_DM, S_LPN, S_TOP, S_INV_DMBH, S_INV_TOPM3, S_BETA, S_LAM, S_KAPPA,
    S_ZP, S_LOPZP, S_DL_LAST, S_FPL, S_CONST, S_LOG_NORM, S_RATE_LOG_NORM,
    S_LPN_D0 = 20,   // d log_pl_norm / d(a, b, mpisn, mbhmax, sigma)          [5]
    S_LN_D0 = 25,    // d log_norm / d(a, b, c, mpisn, mbhmax, sigma, fpl)     [7]
    S_LNV_KAPPA = 32, S_LNV_ZP = 33,
    S_LOG_NSAMP = 34, S_LOG_NDRAW = 35, S_USE_WA = 36, S_DL_FIRST = 37,
    S_EXP_LPN = 38,  // fpl * exp(PISN(mbhmax)) = exp(log_pl_norm)
    S_ZEPS = 39,     // expm1(ZSTEP) = (z_{b+1}-z_b)/(1+z_b), the same for every bin of the log-uniform z grid
    S_C2 = 40,       // 2 exp(log_pl_norm): prefactor of the power-law tail in linear space
    S_LAM2 = 41,     // lam - 2
    S_RATE0 = 42,    // lam - 3 - beta
    S_BAD = 43,      // 1 if theta or any table entry is not finite (outputs are then NaN)
    S_FIXED = 44,    // 1 in fixed-cosmology mode (pop_model): no gradient w.r.t. (h, Om, w)
    S_LOG_C2 = 45,   // log 2 + log_pl_norm: the power-law tail's prefactor, folded into its exponent
    S_POS0 = 46,     // -MIN_BH_MASS * S_INV_DMBH: position on the mbh grid = fma(m, S_INV_DMBH, S_POS0)
};

// ---- per-sample gradient features accumulated by the streaming kernel (DESIGN.md has the algebra)
enum Feat { F_CZ = 0, F_OM, F_W, F_SQ, F_C, F_PA, F_PB, F_PMPISN, F_PMBHMAX, F_PSIGMA, F_GEO, F_T, F_BETA, F_L,
            F_SIG, F_SIGL, F_WA, NFEAT };
static_assert(NFEAT == 17, "17 features");
constexpr int NACC = 2 + NFEAT;   // S, S2, features
constexpr int PART_STRIDE = 24;   // per-record partial: [0] shift m, [1..19] acc, [20] nvalid

// ---- work decomposition of the streaming kernel: the unit is a GROUP of 64 consecutive samples of one event
// (or of the injection set), two samples per lane.  Groups are numbered events-first; every warp of the grid owns
// one contiguous, equally long range of groups and writes one RECORD per event it touches (plus one for the
// injection set), so events never need a block-wide reduction and no warp waits for another.
constexpr int GROUP = 64;
struct Work {
    int64_t g_evt;         // groups per event  = ceil(evt_stride / GROUP)
    int64_t n_evt_groups;  // nobs * g_evt
    int64_t n_groups;      // + groups of the injection set
    int64_t gpw;           // groups per warp (last warps may own fewer / none)
    int64_t evt_stride;    // padded samples per event (multiple of GROUP: the kernel loads whole groups unpredicated)
    int64_t sel_stride;    // padded injections (multiple of GROUP)
    int32_t nobs;
    int32_t nwarps;
};
// event id of a group (the injection set is pseudo-event `nobs`)
__host__ __device__ inline int64_t group_event(const Work& w, const int64_t g) {
    return g < w.n_evt_groups ? g / w.g_evt : (int64_t)w.nobs;
}

constexpr int NCOL = 7;  // dl, m1det, q, log m1det, log q, log1p q, log pdraw
enum Col { C_DL = 0, C_M1D, C_Q, C_LM, C_LQ, C_L1Q, C_LPD };

// Resident data layout: group-blocked SoA.  A set (events, injections) is an array of BLOCKS, one per 64-sample
// group, each holding the 7 columns of its group back to back (7 x 64 doubles = 3584 bytes = 28 lines of 128 B):
//   sample s of group g, column k  ->  base[(g * NCOL + k) * GROUP + s]
// Groups are stored in the order the streaming kernel numbers them, so a warp's range of groups is one contiguous
// stretch of memory: its pointer advances by one block per group, every load is base + immediate offset, and a
// half-warp of loads is still a fully used pair of 128-byte lines per column.
constexpr int BLOCK_DOUBLES = NCOL * GROUP;
struct Columns {
    const double* evt_base;   // nobs * g_evt blocks
    const double* sel_base;   // ceil(nsel / GROUP) blocks
};
__host__ __device__ inline int64_t block_index(const int64_t padded_sample, const int col) {
    return ((padded_sample / GROUP) * NCOL + col) * GROUP + padded_sample % GROUP;
}

// ---- per-rank partial (multi-GPU exchange); doubles
constexpr int PARTIAL_LEN = 128;  // sums [0,64) + copy of the scalar block [64,128)
enum PartialIdx {
    P_LLSUM = 0,      // sum_e [log S_e + m_e]   (no theta-only constants)
    P_NOBS = 1,       // events in this shard
    P_FSUM0 = 2,      // sum_e F_e[k]/S_e, k < 17
    P_NVALID_EVT = 19,
    P_SEL_M = 20,     // injection partial: shift, then acc[19]
    P_SEL_ACC0 = 21,
    P_NVALID_SEL = 40,
    P_NSEL = 41,
    P_NDEAD_EVT = 42, // events without a single finite-weight sample (loglike = -inf)
    P_SCAL0 = 64,
};

// output header (must match include/bump.h)
constexpr int OUT_LOGLIKE = 0, OUT_LOG_MU_SEL = 1, OUT_LOG_MU2 = 2, OUT_NEFF_SEL = 3, OUT_DLOGLIKE = 4,
              OUT_DLOG_MU = 19, OUT_NVALID_EVT = 34, OUT_NVALID_SEL = 35, OUT_NOBS = 36, OUT_NSEL = 37,
              OUT_STATUS = 38, OUT_HEADER = 40;
// OUT_STATUS values (0 = ok): the peer-memory exchange of a multi-rank evaluation failed
constexpr double STATUS_OK = 0.0, STATUS_EXCHANGE_TIMEOUT = 1.0, STATUS_EXCHANGE_POISONED = 2.0;

// ---- optional per-kernel timeline (bump_debug_timeline): first block start / last block end of every kernel of one
// evaluation on the GPU's global nanosecond timer.  tl == nullptr (always, on the normal path) costs one predicate.
enum TimelineSlot { TL_PROLOGUE = 0, TL_STREAM, TL_EPILOGUE, TL_FINALIZE,
                    TL_PRO_ROWS,      // prologue: the 256 PISN-row blocks (start of the first .. end of the last)
                    TL_PRO_COSMO,     // prologue: the 4 cosmology blocks incl. their packing
                    TL_PRO_LAST,      // prologue: the last block's tail (mass records, scalars)
                    TL_STREAM_STAGED, // streaming kernel: first .. last CTA past the table staging
                    TL_EPI_BLOCKS,    // epilogue: per-event / injection phase of all blocks
                    TL_EPI_LAST,      // epilogue: the last block's tail (slot sums, partial, exchange, finalize)
                    TL_STREAM_WARPS,  // streaming kernel: the first .. the last warp to finish its range of groups
                    TL_N };
constexpr int TL_WARP_SLOTS = 4096;   // per-warp end times of the streaming kernel behind the phase slots
#ifdef __CUDACC__
// Programmatic dependent launch (see launch_dependent in bump_lib.cu).  `launch_dependents`: the next kernel on the
// stream may be scheduled once every block of this grid has said so (or exited).  `wait`: returns when the kernel in
// front of this one has completed and its memory is visible; a no-op for a kernel launched in plain stream order.
// Measured on B200 (profiles/r02_ab_experiments.md, GWTC-3 shape, same box):
//   stream kernel -> epilogue, dependents released when the stream kernel's blocks exit (default): -0.7 us
//   ... released when the first warp of every block is done (BUMP_PDL_EPI_TRIGGER=1):              -0.6 us
//   ... released at kernel start (=2): -0.5 us or nothing where the grid leaves SMs idle for every epilogue block
//       (GWTC-3 shape), +4 us where it does not (8-GPU shard: the waiting blocks queue up on the one idle SM)
//   prologue -> stream kernel as well (BUMP_PDL_STREAM): nothing on top (the stream kernel needs a whole SM's shared
//       memory, so it cannot start beside a prologue block), and with dependents released at the prologue's start
//       concurrent contexts returned WRONG results (the stream kernel's constant-bank reads are only ordered
//       against the prologue's writes by a launch that follows the prologue's completion).  Off, and not to be used.
#ifdef BUMP_PDL_STREAM
constexpr bool PDL_STREAM = true;      // prologue -> stream kernel (measurement option, see above)
#else
constexpr bool PDL_STREAM = false;
#endif
#ifdef BUMP_NO_PDL
constexpr bool PDL_EPILOGUE = false;
#else
constexpr bool PDL_EPILOGUE = true;    // stream kernel -> epilogue
#endif
#ifndef BUMP_PDL_EPI_TRIGGER
#define BUMP_PDL_EPI_TRIGGER 0   // 0: when the stream kernel's blocks exit, 1: when their first warp is done, 2: at their start
#endif
#ifndef BUMP_PDL_STREAM_TRIGGER
#define BUMP_PDL_STREAM_TRIGGER 0   // 0: when the prologue's blocks exit, 2: at their start
#endif
template <bool ON>
__device__ __forceinline__ void pdl_launch_dependents() {
    if (ON) asm volatile("griddepcontrol.launch_dependents;");
}
template <bool ON>
__device__ __forceinline__ void pdl_wait() {
    if (ON) asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void timeline_begin(unsigned long long* tl, const int k) {
    if (tl != nullptr && threadIdx.x == 0) atomicMin(tl + 2 * k, global_ns());
}
__device__ __forceinline__ void timeline_end(unsigned long long* tl, const int k) {
    if (tl != nullptr && threadIdx.x == 0) atomicMax(tl + 2 * k + 1, global_ns());
}
#endif

}  // namespace bump

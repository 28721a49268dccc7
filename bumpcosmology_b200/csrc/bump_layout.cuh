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
constexpr double LN2 = 0.69314718055994530942;
constexpr double HALF_LOG_2PI = 0.91893853320467274178;     // log(sqrt(2 pi))
constexpr double FOUR_PI = 12.566370614359172954;

constexpr int NTHETA = 14;
constexpr int NTHETA_MAX = 15;
enum ThetaIdx { T_H = 0, T_OM, T_W, T_A, T_B, T_C, T_MPISN, T_MBHMAX, T_SIGMA, T_FPL, T_BETA, T_LAM, T_KAPPA,
                T_ZP, T_WA };

// ---- theta-dependent table blob: built by the prologue kernel in global memory, bulk-copied (TMA) to
// shared memory by every CTA of the streaming kernel.  Offsets in doubles.
constexpr int NSCAL = 64;
constexpr int NCREC = 10;  // cosmology per-bin records {f0, f1-f0}
constexpr int NMREC = 6;   // mass per-bin records
constexpr int OFF_SCAL = 0;
constexpr int OFF_COS = OFF_SCAL + NSCAL;           // double2 cos[NCREC][NZ]
constexpr int OFF_DLK = OFF_COS + NCREC * NZ * 2;   // double dlk[NZ]     (search keys)
constexpr int OFF_MASS = OFF_DLK + NZ;              // double2 mass[NMREC][NM]
constexpr int BLOB_DOUBLES = OFF_MASS + NMREC * NM * 2;
constexpr int BLOB_BYTES = BLOB_DOUBLES * 8;
static_assert(BLOB_BYTES % 16 == 0, "bulk copies need 16-byte multiples");

// cosmology records, bin b = [knot b, knot b+1]
enum CosRec { CR_DL = 0,   // {dl_b, 1/(dl_{b+1}-dl_b)}
              CR_DVC,      // {dvc_b, dvc_{b+1}-dvc_b}
              CR_DDL,      // {ddl_b, ...}
              CR_DL_OM, CR_DL_W, CR_DVC_OM, CR_DVC_W, CR_DDL_OM, CR_DDL_W,   // tangent tables, same form
              CR_Z };      // {1/(1+z_b), log(1+z_b)}     theta-independent
// mass records, bin b of the mbh grid
enum MassRec { MR_G = 0, MR_GA, MR_GB, MR_GMPISN, MR_GMBHMAX, MR_GSIGMA };

// scalars
enum Scal {
    S_H = 0, S_INV_H, S_C, S_M, S_LOG_M, S_INV_DM, S_LPN, S_TOP, S_INV_DMBH, S_INV_TOPM3, S_BETA, S_LAM, S_KAPPA,
    S_ZP, S_LOPZP, S_DL_LAST, S_FPL, S_CONST, S_LOG_NORM, S_RATE_LOG_NORM,
    S_LPN_D0 = 20,   // d log_pl_norm / d(a, b, mpisn, mbhmax, sigma)          [5]
    S_LN_D0 = 25,    // d log_norm / d(a, b, c, mpisn, mbhmax, sigma, fpl)     [7]
    S_LNV_KAPPA = 32, S_LNV_ZP = 33,
    S_LOG_NSAMP = 34, S_LOG_NDRAW = 35, S_NOBS_LOCAL = 36,
};

// ---- per-sample gradient features accumulated by the streaming kernel (see DESIGN.md for the algebra)
enum Feat { F_CZ = 0, F_OM, F_W, F_SQ, F_C, F_PA, F_PB, F_PMPISN, F_PMBHMAX, F_PSIGMA, F_GEO, F_T, F_BETA, F_L,
            F_SIG, F_SIGL, NFEAT };
static_assert(NFEAT == 16, "16 features");
constexpr int NACC = 2 + NFEAT;   // S, S2, features
constexpr int PART_STRIDE = 20;   // per-tile partial: m, acc[18], nvalid

// ---- tiles
struct Tile {
    int64_t off;   // first sample (index into the padded column arrays of its set)
    int32_t count; // samples in the tile (even; may include sentinel padding)
    int32_t set;   // 0 = events, 1 = injections
};

constexpr int NCOL = 7;  // dl, m1det, q, log m1det, log q, log1p q, log pdraw
enum Col { C_DL = 0, C_M1D, C_Q, C_LM, C_LQ, C_L1Q, C_LPD };

struct Columns {
    const double* evt[NCOL];
    const double* sel[NCOL];
};

// ---- per-rank partial (multi-GPU exchange); doubles
constexpr int PARTIAL_SUMS = 64;
constexpr int PARTIAL_LEN = 128;  // sums + copy of the scalars (for the host-side merge)
enum PartialIdx {
    P_LLSUM = 0,      // sum_e [log S_e + m_e]   (no constants)
    P_NOBS = 1,       // events in this shard
    P_FSUM0 = 2,      // sum_e F_e[k]/S_e, k < 16
    P_NVALID_EVT = 18,
    P_SEL_M = 19,     // injection partial: running max, then acc[18]
    P_SEL_ACC0 = 20,
    P_NVALID_SEL = 38,
    P_NSEL = 39,
};

// output header (must match include/bump.h)
constexpr int OUT_LOGLIKE = 0, OUT_LOG_MU_SEL = 1, OUT_LOG_MU2 = 2, OUT_NEFF_SEL = 3, OUT_DLOGLIKE = 4,
              OUT_DLOG_MU = 19, OUT_NVALID_EVT = 34, OUT_NVALID_SEL = 35, OUT_HEADER = 40;

}  // namespace bump

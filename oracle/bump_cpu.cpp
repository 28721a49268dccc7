// oracle/bump_cpu.cpp — fused single-pass C++/OpenMP port of the hot path.  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.
//
// A CPU implementation of the same function the CUDA library computes (theta -> loglike, log_mu_sel, log_mu2,
// neff_sel, neff[nobs], d loglike/d theta, d log_mu_sel/d theta), written the way a fusing CPU compiler would run
// the reference: one pass over the samples, everything in LOG space with libm exp/log/log1p exactly where
// /root/reference/src/scripts/intensity_models.py has them (:45-54 turn-on, :140-151 logaddexp, :170-173 rate,
// :202-210 joint density, :378-394 weights and logsumexps, :401 Neff), theta-independent logs hoisted, OpenMP over
// events and injection chunks.  It serves two purposes:
//   * bench.py's CPU arm (`cpu_baseline`, `--impl reference`): a stronger stand-in for "the reference's JAX on the
//     host cores" than the eager torch oracle (SURVEY.md section 8d asks for both and for the faster as denominator);
//   * a third, independently written evaluation (log space, libm, binary search) checked against the goldens in
//     tests/test_cpu_port.py.
// Only tests/, __graft_entry__ and bench.py's CPU legs may load it.  Build: `make -C oracle` -> oracle/_build/.
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "cpu_dual.h"   // the oracle's own forward-mode derivatives (no source shared with the CUDA library)

template <int N>
using Dual = cpuad::Fwd<N>;

namespace {

// constants: file:line in intensity_models.py
constexpr double MBH_MIN = 5.0, MTR = 20.0, TURNON_WIDTH = 0.05, MIN_BH_MASS = 3.0, MIN_CO_MASS = 1.0;   // :13,41,45,97,98
constexpr double MREF = 30.0, ZMAX = 100.0, C_H100_GPC = 2.99792;                                          // :129,220,239
constexpr int NM = 256, NZ = 1024;                                                                         // :92,221
constexpr double LN2 = 0.69314718055994530942, HALF_LOG_2PI = 0.91893853320467274178, FOUR_PI = 12.566370614359172954;
constexpr double LOG60 = 4.0943445622221004;

enum { T_H, T_OM, T_W, T_A, T_B, T_C, T_MPISN, T_MBHMAX, T_SIGMA, T_FPL, T_BETA, T_LAM, T_KAPPA, T_ZP, NTH };
enum { F_CZ, F_OM, F_W, F_SQ, F_C, F_PA, F_PB, F_PMPISN, F_PMBHMAX, F_PSIGMA, F_GEO, F_T, F_BETA, F_L, F_SIG, F_SIGL, NF };
constexpr int OUT_LOGLIKE = 0, OUT_LOG_MU = 1, OUT_LOG_MU2 = 2, OUT_NEFF_SEL = 3, OUT_DLL = 4, OUT_DMU = 19, OUT_HEADER = 40;

struct Tables {
    double z[NZ], dl[NZ], ddl[NZ], dvc[NZ];
    double t_dl[2][NZ], t_ddl[2][NZ], t_dvc[2][NZ];   // d/dOm, d/dw
    double G[NM], tG[5][NM];                          // log_dN_grid and d/d(a, b, mpisn, mbhmax, sigma)
    double top, inv_dmbh, M, logM, inv_dm, c, lpn, beta, lam, kappa, lopzp, cst, log_norm, lnV;
    double lpn_d[5], ln_d[7], lnv_kappa, lnv_zp, h, fpl, zp;
};

void build_tables(const double* th, Tables& T) {
    // ---- flat wCDM tables with tangents (:229-235, utils.py:3-8)
    {
        typedef Dual<2> D;
        const D Om = D::seed(th[T_OM], 0), w = D::seed(th[T_W], 1);
        const double dH = C_H100_GPC / th[T_H];
        const double step = log1p(ZMAX) / (NZ - 1);
        std::vector<D> iE(NZ);
        for (int k = 0; k < NZ; ++k) {
            const double lz = (k == NZ - 1) ? log1p(ZMAX) : k * step;
            T.z[k] = expm1(lz);
            const double opz = 1.0 + T.z[k];
            D de = cpuad::exp((3.0 * (1.0 + w)) * log(opz));
            D E = cpuad::sqrt(Om * (opz * opz * opz) + (1.0 - Om) * de);
            iE[k] = 1.0 / E;
        }
        D C(0.0);
        for (int k = 0; k < NZ; ++k) {
            if (k > 0) C = C + (0.5 * (T.z[k] - T.z[k - 1])) * (iE[k - 1] + iE[k]);
            const double opz = 1.0 + T.z[k];
            D dc = dH * C, dl = dc * opz, ddl = dc + (dH * opz) * iE[k], dvc = (FOUR_PI * dH) * (cpuad::square(dc) * iE[k]);
            T.dl[k] = dl.v, T.ddl[k] = ddl.v, T.dvc[k] = dvc.v;
            for (int c = 0; c < 2; ++c) T.t_dl[c][k] = dl.d[c], T.t_ddl[c][k] = ddl.d[c], T.t_dvc[c][k] = dvc.d[c];
        }
    }
    // ---- PISN table (:96-108), one row per iteration
    typedef Dual<5> D5;
    const D5 a = D5::seed(th[T_A], 0), b = D5::seed(th[T_B], 1), mpisn = D5::seed(th[T_MPISN], 2),
             M = D5::seed(th[T_MBHMAX], 3), sg = D5::seed(th[T_SIGMA], 4);
    const D5 top = M + 7.0 * sg, mcomax = 2.0 * M - mpisn, mco_top = mcomax + cpuad::sqrt(4.0 * M * (M - mpisn));
    const D5 alpha = 1.0 / (4.0 * (mpisn - M));
    std::vector<D5> mco(NM), ell(NM), mu(NM);
    for (int j = 0; j < NM; ++j) {
        const double sj = (double)j / (NM - 1);
        mco[j] = (j == NM - 1) ? mco_top : (MIN_CO_MASS * (1.0 - sj) + mco_top * sj);
        mu[j] = (mco[j].v < mpisn.v) ? mco[j] : (M + alpha * cpuad::square(mco[j] - mcomax));
        D5 lx = cpuad::log(mco[j] / MTR);
        ell[j] = (mco[j].v < MTR) ? (-a * lx) : (-b * lx);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < NM; ++i) {
        const double si = (double)i / (NM - 1);
        const D5 mbh = (i == NM - 1) ? top : (MIN_BH_MASS * (1.0 - si) + top * si);
        D5 lw[NM];
        for (int j = 0; j < NM; ++j) {
            D5 u = (mbh - mu[j]) / sg;
            lw[j] = ell[j] - 0.5 * cpuad::square(u) - HALF_LOG_2PI - cpuad::log(sg);
        }
        D5 term[NM - 1];
        double mx = -INFINITY;
        for (int j = 0; j < NM - 1; ++j) {
            term[j] = cpuad::logaddexp(lw[j + 1], lw[j]) + cpuad::log(mco[j + 1] - mco[j]) - LN2;
            mx = fmax(mx, term[j].v);
        }
        double s = 0.0, sd[5] = {0, 0, 0, 0, 0};
        for (int j = 0; j < NM - 1; ++j) {
            const double e = exp(term[j].v - mx);
            s += e;
            for (int k = 0; k < 5; ++k) sd[k] += e * term[j].d[k];
        }
        T.G[i] = mx + log(s);
        for (int k = 0; k < 5; ++k) T.tG[k][i] = sd[k] / s;
    }
    // ---- scalars (:134-138, :167-168)
    typedef Dual<7> D7;   // a, b, c, mpisn, mbhmax, sigma, fpl
    const int map5[5] = {0, 1, 3, 4, 5};
    const D7 c7 = D7::seed(th[T_C], 2), M7 = D7::seed(th[T_MBHMAX], 4), sg7 = D7::seed(th[T_SIGMA], 5),
             fpl7 = D7::seed(th[T_FPL], 6);
    const D7 top7 = M7 + 7.0 * sg7;
    auto knot = [&](int k) -> D7 {
        const double s = (double)k / (NM - 1);
        return (k == NM - 1) ? top7 : (MIN_BH_MASS * (1.0 - s) + top7 * s);
    };
    auto Gk = [&](int k) -> D7 {
        D7 g(T.G[k]);
        for (int q = 0; q < 5; ++q) g.d[map5[q]] = T.tG[q][k];
        return g;
    };
    auto pisn = [&](const D7& m) -> D7 {   // jnp.interp, differentiable in x, xp, fp
        int i = 1;
        while (i < NM - 1 && knot(i).v <= m.v) ++i;
        D7 x0 = knot(i - 1), x1 = knot(i), f0 = Gk(i - 1), f1 = Gk(i);
        D7 f = f0 + ((m - x0) / (x1 - x0)) * (f1 - f0);
        if (m.v < knot(0).v) f = Gk(0);
        if (m.v > top7.v) f = Gk(NM - 1);
        return f;
    };
    const D7 lpn = cpuad::log(fpl7) + pisn(M7);
    const D7 mref(MREF);
    const D7 P = (MREF <= MIN_BH_MASS || MREF >= top7.v) ? D7(-INFINITY) : pisn(mref);
    const D7 turn = LN2 - cpuad::log1p(cpuad::exp(-(mref - M7) / (M7 * TURNON_WIDTH)));
    const D7 Q = -c7 * cpuad::log(mref / M7) + lpn + turn;
    const D7 ln = -(cpuad::logaddexp(P, Q) + log(MREF));
    const double kappa = th[T_KAPPA], zp = th[T_ZP], lopzp = log1p(zp);
    const double r0 = exp(-kappa * lopzp), sig0 = r0 / (1.0 + r0);
    T.top = top7.v, T.inv_dmbh = (NM - 1) / (top7.v - MIN_BH_MASS), T.M = M7.v, T.logM = log(M7.v);
    T.inv_dm = 1.0 / (M7.v * TURNON_WIDTH), T.c = th[T_C], T.lpn = lpn.v, T.beta = th[T_BETA], T.lam = th[T_LAM];
    T.kappa = kappa, T.lopzp = lopzp, T.log_norm = ln.v, T.lnV = log1p(r0), T.h = th[T_H], T.fpl = th[T_FPL], T.zp = zp;
    T.cst = 2.0 * ln.v + T.lnV - th[T_BETA] * LOG60;
    for (int q = 0; q < 5; ++q) T.lpn_d[q] = lpn.d[map5[q]];
    for (int q = 0; q < 7; ++q) T.ln_d[q] = ln.d[q];
    T.lnv_kappa = -sig0 * lopzp, T.lnv_zp = -sig0 * kappa / (1.0 + zp);
}

// log dN/dm without its normalisation (:140-151) and the softmax-weighted gradient features of one mass
inline double mass_term(const Tables& T, const double m, const double lm, double* f, double& mdA) {
    const double y = (m - T.M) * T.inv_dm;
    const double e = exp(-y);
    const double turn = LN2 - log1p(e);              // :52-54
    const double sg = e / (1.0 + e);
    const double lrel = lm - T.logM;
    const double Q = -T.c * lrel + T.lpn + turn;     // :147
    const double pos = (m - MIN_BH_MASS) * T.inv_dmbh;
    const int b = std::min(std::max((int)floor(pos), 0), NM - 2);
    const double u = pos - b;
    const double slope = (T.G[b + 1] - T.G[b]) * T.inv_dmbh;
    const double P = (m < T.top) ? T.G[b] + u * (T.G[b + 1] - T.G[b]) : -INFINITY;   // :144-145
    const double mx = fmax(P, Q);
    const double eP = exp(P - mx), eQ = exp(Q - mx), s = eP + eQ;
    const double sP = eP / s, sQ = eQ / s;
    mdA = sP * slope * m + sQ * (sg * T.inv_dm * m - T.c);
    f[F_SQ] += sQ;
    f[F_C] += sQ * lrel;
    for (int k = 0; k < 5; ++k) f[F_PA + k] += sP * (T.tG[k][b] + u * (T.tG[k][b + 1] - T.tG[k][b]));
    f[F_GEO] += sP * slope * (m - MIN_BH_MASS);
    f[F_T] += sQ * sg * m;
    return mx + log(s);
}

// log weight of one sample (:378-381) and its features; returns -inf for zero weight
inline double sample(const Tables& T, const double x, const double m1d, const double q, const double lm,
                     const double lq, const double l1q, const double lpd, double* f) {
    // z_of_dL = jnp.interp(dl, dlinterp, zinterp): searchsorted(side='right') clipped to [1, n-1]
    int i = (int)(std::upper_bound(T.dl, T.dl + NZ, x) - T.dl);
    i = std::min(std::max(i, 1), NZ - 1);
    const int b = i - 1;
    const bool beyond = x > T.dl[NZ - 1];
    const double ddlb = T.dl[b + 1] - T.dl[b];
    const double t = beyond ? 1.0 : (x - T.dl[b]) / ddlb;
    const double idl = beyond ? 0.0 : 1.0 / ddlb;
    const double z = T.z[b] + t * (T.z[b + 1] - T.z[b]);
    const double L = log1p(z);
    const double m1 = m1d / (1.0 + z), m2 = q * m1;
    if (m1 < MBH_MIN || m2 < MBH_MIN) return -INFINITY;   // :149
    const double dvc = T.dvc[b] + t * (T.dvc[b + 1] - T.dvc[b]);
    const double ddl = T.ddl[b] + t * (T.ddl[b + 1] - T.ddl[b]);
    if (!(dvc > 0.0)) return -INFINITY;
    const double lm1 = lm - L, lm2 = lm1 + lq;
    for (int k = 0; k < NF; ++k) f[k] = 0.0;
    double mdA1, mdA2;
    const double A1 = mass_term(T, m1, lm1, f, mdA1), A2 = mass_term(T, m2, lm2, f, mdA2);
    const double r = exp(T.kappa * (L - T.lopzp));
    const double sig = r / (1.0 + r);
    const double V0 = T.lam * L - log1p(r);            // :173
    const double pair = lm1 + l1q;
    const double w = A1 + A2 + T.beta * pair + lm1 + V0 - 2.0 * L + log(dvc) - log(ddl) - lpd;   // :210, :381
    // d w / d t at fixed tables; dL/dt = (z_{b+1} - z_b) / (1 + z)
    const double lt = (T.z[b + 1] - T.z[b]) / (1.0 + z);
    const double Wt = lt * (T.lam - T.kappa * sig - 3.0 - T.beta - mdA1 - mdA2) + (T.dvc[b + 1] - T.dvc[b]) / dvc -
                      (T.ddl[b + 1] - T.ddl[b]) / ddl;
    const double Wx = Wt * idl;
    f[F_CZ] = Wx * x;
    for (int c = 0; c < 2; ++c) {
        const double tdl = T.t_dl[c][b] + t * (T.t_dl[c][b + 1] - T.t_dl[c][b]);
        const double tdvc = T.t_dvc[c][b] + t * (T.t_dvc[c][b + 1] - T.t_dvc[c][b]);
        const double tddl = T.t_ddl[c][b] + t * (T.t_ddl[c][b + 1] - T.t_ddl[c][b]);
        f[F_OM + c] = -Wx * tdl + tdvc / dvc - tddl / ddl;
    }
    f[F_BETA] = pair, f[F_L] = L, f[F_SIG] = sig, f[F_SIGL] = sig * L;
    return w;
}

struct Lse {   // max-shifted accumulator of (S, S2, features)
    double m = -INFINITY, S = 0, S2 = 0, F[NF] = {0};
    void add(const double w, const double* f) {
        if (w == -INFINITY) return;
        if (w > m) {
            const double s = (m == -INFINITY) ? 0.0 : exp(m - w);
            S *= s, S2 *= s * s;
            for (int k = 0; k < NF; ++k) F[k] *= s;
            m = w;
        }
        const double p = exp(w - m);
        S += p, S2 += p * p;
        for (int k = 0; k < NF; ++k) F[k] += p * f[k];
    }
    void merge(const Lse& o) {
        if (o.m == -INFINITY) return;
        const double mx = fmax(m, o.m), a = (m == -INFINITY) ? 0.0 : exp(m - mx), b = exp(o.m - mx);
        S = S * a + o.S * b, S2 = S2 * a * a + o.S2 * b * b;
        for (int k = 0; k < NF; ++k) F[k] = F[k] * a + o.F[k] * b;
        m = mx;
    }
};

void grad_from_features(const Tables& T, const double* phi, const double n, double* g) {
    const double sq = phi[F_SQ], geo = phi[F_GEO] / (T.top - MIN_BH_MASS);
    g[T_H] = (phi[F_CZ] - 2.0 * n) / T.h;
    g[T_OM] = phi[F_OM], g[T_W] = phi[F_W];
    g[T_A] = phi[F_PA] + T.lpn_d[0] * sq + 2.0 * n * T.ln_d[0];
    g[T_B] = phi[F_PB] + T.lpn_d[1] * sq + 2.0 * n * T.ln_d[1];
    g[T_C] = -phi[F_C] + 2.0 * n * T.ln_d[2];
    g[T_MPISN] = phi[F_PMPISN] + T.lpn_d[2] * sq + 2.0 * n * T.ln_d[3];
    g[T_MBHMAX] = phi[F_PMBHMAX] - geo + sq * (T.c / T.M + T.lpn_d[3]) - phi[F_T] * T.inv_dm / T.M + 2.0 * n * T.ln_d[4];
    g[T_SIGMA] = phi[F_PSIGMA] - 7.0 * geo + T.lpn_d[4] * sq + 2.0 * n * T.ln_d[5];
    g[T_FPL] = sq / T.fpl + 2.0 * n * T.ln_d[6];
    g[T_BETA] = phi[F_BETA] - n * LOG60;
    g[T_LAM] = phi[F_L];
    g[T_KAPPA] = -(phi[F_SIGL] - T.lopzp * phi[F_SIG]) + n * T.lnv_kappa;
    g[T_ZP] = phi[F_SIG] * T.kappa / (1.0 + T.zp) + n * T.lnv_zp;
}

struct Catalog {
    int64_t nobs, nsamp, nsel;
    double ndraw;
    std::vector<double> e[7], s[7];   // dl, m1d, q, log m1d, log q, log1p q, log pdraw
};

void fill(std::vector<double>* c, int64_t n, const double* m1d, const double* q, const double* dl, const double* pd) {
    for (int k = 0; k < 7; ++k) c[k].resize(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        c[0][i] = dl[i], c[1][i] = m1d[i], c[2][i] = q[i];
        c[3][i] = log(m1d[i]), c[4][i] = log(q[i]), c[5][i] = log1p(q[i]), c[6][i] = log(pd[i]);
    }
}

}  // namespace

extern "C" {

void* bcpu_create(int64_t nobs, int64_t nsamp, const double* m1d, const double* q, const double* dl, const double* pd,
                  int64_t nsel, const double* sm1d, const double* sq, const double* sdl, const double* spd, double ndraw) {
    Catalog* c = new Catalog();
    c->nobs = nobs, c->nsamp = nsamp, c->nsel = nsel, c->ndraw = ndraw;
    fill(c->e, nobs * nsamp, m1d, q, dl, pd);
    fill(c->s, nsel, sm1d, sq, sdl, spd);
    return c;
}

void bcpu_destroy(void* h) { delete static_cast<Catalog*>(h); }

int bcpu_max_threads(void) { return omp_get_max_threads(); }

// out: OUT_HEADER + nobs doubles, same layout as include/bump.h
int bcpu_eval(void* h, const double* theta, double* out, int nthreads) {
    const Catalog& c = *static_cast<Catalog*>(h);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    std::vector<Tables> holder(1);   // ~110 KB: on the heap, shared by the OpenMP team
    Tables& T = holder[0];
    build_tables(theta, T);
    for (int k = 0; k < OUT_HEADER; ++k) out[k] = 0.0;
    double llsum = 0.0, phi[NF] = {0};
    int dead = 0;
    double* neff = out + OUT_HEADER;
#pragma omp parallel
    {
        double f[NF], lphi[NF] = {0}, lll = 0.0;
        int ldead = 0;
#pragma omp for schedule(dynamic, 4) nowait
        for (int64_t e = 0; e < c.nobs; ++e) {
            Lse A;
            const int64_t o = e * c.nsamp;
            for (int64_t j = 0; j < c.nsamp; ++j) {
                const int64_t i = o + j;
                const double w = sample(T, c.e[0][i], c.e[1][i], c.e[2][i], c.e[3][i], c.e[4][i], c.e[5][i], c.e[6][i], f);
                A.add(w, f);
            }
            if (A.m == -INFINITY) {
                ++ldead;
                neff[e] = NAN;
                continue;
            }
            lll += A.m + log(A.S);                    // logsumexp (:382)
            neff[e] = A.S * A.S / A.S2;               // :401
            for (int k = 0; k < NF; ++k) lphi[k] += A.F[k] / A.S;
        }
#pragma omp critical
        {
            llsum += lll, dead += ldead;
            for (int k = 0; k < NF; ++k) phi[k] += lphi[k];
        }
    }
    const double cst = T.cst;
    out[OUT_LOGLIKE] = dead ? -INFINITY : llsum + c.nobs * (cst - log((double)c.nsamp));
    grad_from_features(T, phi, (double)c.nobs, out + OUT_DLL);
    // injections (:385-394)
    Lse S;
#pragma omp parallel
    {
        Lse A;
        double f[NF];
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < c.nsel; ++i)
            A.add(sample(T, c.s[0][i], c.s[1][i], c.s[2][i], c.s[3][i], c.s[4][i], c.s[5][i], c.s[6][i], f), f);
#pragma omp critical
        S.merge(A);
    }
    const double lnd = log(c.ndraw);
    const double log_mu = S.m + log(S.S) + cst - lnd, log_mu2 = 2.0 * S.m + log(S.S2) + 2.0 * cst - 2.0 * lnd;
    const double log_s2 = log_mu2 + log1p(-exp(2.0 * log_mu - lnd - log_mu2));
    out[OUT_LOG_MU] = log_mu, out[OUT_LOG_MU2] = log_mu2, out[OUT_NEFF_SEL] = exp(2.0 * log_mu - log_s2);
    double phis[NF];
    for (int k = 0; k < NF; ++k) phis[k] = S.F[k] / S.S;
    grad_from_features(T, phis, 1.0, out + OUT_DMU);
    out[36] = (double)c.nobs, out[37] = (double)c.nsel;
    return 0;
}

}  // extern "C"

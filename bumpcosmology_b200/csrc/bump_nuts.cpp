// Host NUTS driver in C++ (no GIL, no interpreter in the leapfrog loop): one chain per call, chains run in parallel
// from concurrent host threads, each on its own context (contexts of one device evaluate concurrently on different
// constant-bank slots, bump_lib.cu).
//
// Replaces what numpyro runs for the reference (third-party, unpinned; /root/reference/src/scripts/run_cosmo_fit.py:
// 17-19 seed / chain configuration, :45-49 `NUTS(pop_cosmo_model, dense_mass=True)`, `MCMC(num_warmup=1000,
// num_samples=1000, num_chains=4)`): multinomial NUTS with the generalised U-turn criterion, max tree depth 10,
// divergence threshold 1000, dual-averaging step-size adaptation to a target acceptance of 0.8, Stan-style windowed
// adaptation of a dense mass matrix.  It is a statement-for-statement port of bumpcosmology_b200/nuts.py (the Python
// driver stays as the readable specification and as the cross-check in tests/test_nuts_native.py); priors and
// transforms follow bumpcosmology_b200/priors.py (intensity_models.py:281-311,398).  Not bit-compatible with either
// random stream (mt19937_64 here); acceptance is statistical.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "../../include/bump.h"

extern "C" int bump_set_error(int code, const char* msg);   // bump_lib.cu: thread-local message + code

namespace {

constexpr int MAXD = 32;
constexpr int NREC = BUMP_NUTS_NDET;
constexpr double MAX_DELTA_H = 1000.0;
constexpr double INF = std::numeric_limits<double>::infinity();

// ------------------------------------------------------------------ potentials
struct Potential {
    int dim = 0;
    long n_evals = 0;
    virtual ~Potential() {}
    // U(u) and dU/du; rec[NREC] = light record of the evaluation (deterministic sites).  Non-finite U: g is zeroed.
    virtual double eval(const double* u, double* g, double* rec) = 0;
    virtual void constrain(const double* u, double* x) const { memcpy(x, u, sizeof(double) * dim); }
    virtual int error() const { return 0; }
};

struct CallbackPotential : Potential {
    bump_potential_cb f;
    void* user;
    CallbackPotential(bump_potential_cb f_, void* user_, int d) : f(f_), user(user_) { dim = d; }
    double eval(const double* u, double* g, double* rec) override {
        ++n_evals;
        for (int k = 0; k < NREC; ++k) rec[k] = 0.0;
        double U = f(user, u, g);
        bool ok = std::isfinite(U);
        for (int i = 0; i < dim && ok; ++i) ok = std::isfinite(g[i]);
        if (!ok) {
            for (int i = 0; i < dim; ++i) g[i] = 0.0;
            return INF;
        }
        return U;
    }
};

// The 15 sample sites of pop_cosmo_model in declaration order (priors.py SITES; intensity_models.py:282-309,398)
struct Site {
    char kind;   // 't' truncated normal, 'n' normal, 'u' uniform
    double a, b, lo, hi, log_z;
};
double Phi(const double x) { return 0.5 * erfc(-x / sqrt(2.0)); }
Site make_site(const char kind, const double a, const double b, const double lo = -INF, const double hi = INF) {
    Site s{kind, a, b, lo, hi, 0.0};
    if (kind == 'u') s.lo = a, s.hi = b;
    if (kind == 't') {
        const double pa = std::isfinite(lo) ? Phi((lo - a) / b) : 0.0, pb = std::isfinite(hi) ? Phi((hi - a) / b) : 1.0;
        s.log_z = log(pb - pa);
    }
    return s;
}
constexpr int NSITES = 15;
const Site* sites() {
    static const Site S[NSITES] = {
        make_site('t', 0.7, 0.2, 0.35, 1.4),            // h        :306
        make_site('t', 0.3, 0.15, 0.0, 1.0),            // Om       :307
        make_site('t', -1.0, 0.25, -1.5, -0.5),         // w        :308
        make_site('t', 2.35, 2.0, -1.65, 6.35),         // a        :282
        make_site('t', 1.9, 2.0, -2.1, 5.9),            // b        :283
        make_site('t', 4.0, 2.0, 0.0, 8.0),             // c        :284
        make_site('t', 35.0, 5.0, 20.0, 50.0),          // mpisn    :286
        make_site('t', 5.0, 2.0, 0.5, 11.0),            // dmbhmax  :287
        make_site('t', 2.0, 2.0, 1.0),                  // sigma    :289
        make_site('n', 0.0, 2.0),                       // beta     :291
        make_site('u', log(1e-3), log(0.5)),            // log_fpl  :293
        make_site('t', 2.7, 2.0, -1.3, 6.7),            // lam      :299
        make_site('t', 5.6 - 2.7, 2.0, 1.0, 9.6 - 2.7), // dkappa   :300
        make_site('t', 1.9, 1.0, 0.0, 3.9),             // zp       :302
        make_site('n', 0.0, 1.0),                       // R_unit   :398
    };
    return S;
}

// unconstrained u -> x, dx/du, d log|dx/du| / du, d log prior / dx; returns log prior(x) + log|dx/du|
double site_terms(const Site& s, const double u, double& x, double& dx, double& dlj, double& glp) {
    const double LOG_SQRT_2PI = 0.91893853320467274178;
    double total = 0.0;
    if (std::isfinite(s.lo) && std::isfinite(s.hi)) {   // interval: lo + (hi - lo) sigmoid(u)
        double sg, lsig2;
        if (u >= 0) {
            const double e = exp(-u);
            sg = 1.0 / (1.0 + e);
            lsig2 = -u - 2.0 * log1p(e);
        } else {
            const double e = exp(u);
            sg = e / (1.0 + e);
            lsig2 = u - 2.0 * log1p(e);
        }
        const double w = s.hi - s.lo;
        x = s.lo + w * sg;
        dx = w * sg * (1.0 - sg);
        total += log(w) + lsig2;
        dlj = 1.0 - 2.0 * sg;
    } else if (std::isfinite(s.lo)) {   // greater_than(lo): lo + exp(u)
        const double e = exp(u);
        x = s.lo + e, dx = e, dlj = 1.0;
        total += u;
    } else {
        x = u, dx = 1.0, dlj = 0.0;
    }
    if (s.kind == 'u') {
        total -= log(s.hi - s.lo);
        glp = 0.0;
    } else {
        const double z = (x - s.a) / s.b;
        total += -0.5 * z * z - LOG_SQRT_2PI - log(s.b) - (s.kind == 't' ? s.log_z : 0.0);
        glp = -z / s.b;
    }
    return total;
}

// U(u) = -[log prior(x(u)) + log|dx/du| + loglike + selfactor] of pop_cosmo_model (intensity_models.py:357-401),
// one library evaluation per call (bumpcosmology_b200/intensity_models.py `potential`)
struct ModelPotential : Potential {
    bump_ctx* ctx;
    std::vector<double> out;
    int err = 0;
    explicit ModelPotential(bump_ctx* c) : ctx(c), out((size_t)bump_out_len(c)) { dim = NSITES; }
    int error() const override { return err; }
    void constrain(const double* u, double* x) const override {
        double dx, dlj, glp;
        for (int i = 0; i < NSITES; ++i) site_terms(sites()[i], u[i], x[i], dx, dlj, glp);
    }
    double eval(const double* u, double* g, double* rec) override {
        double x[NSITES], dx[NSITES], dlj[NSITES], glp[NSITES], lpj = 0.0;
        for (int i = 0; i < NSITES; ++i) lpj += site_terms(sites()[i], u[i], x[i], dx[i], dlj[i], glp[i]);
        const double fpl = exp(x[10]);
        // derived kernel parameters (:288, :294, :301): h Om w a b c mpisn mbhmax sigma fpl beta lam kappa zp
        const double theta[BUMP_NTHETA] = {x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[6] + x[7], x[8], fpl, x[9],
                                           x[11], x[11] + x[12], x[13]};
        ++n_evals;
        if (int r = bump_eval(ctx, theta, out.data())) {
            err = r;
            for (int i = 0; i < NSITES; ++i) g[i] = 0.0;
            return INF;
        }
        const double nobs = out[BUMP_OUT_NOBS], loglike = out[BUMP_OUT_LOGLIKE], log_mu = out[BUMP_OUT_LOG_MU_SEL];
        const double logl = loglike - nobs * log_mu;
        const double mu = std::isfinite(log_mu) ? exp(log_mu) : NAN;
        double neff_min = NAN;
        for (size_t k = BUMP_OUT_HEADER; k < out.size(); ++k)
            if (!(neff_min <= out[k])) neff_min = out[k];   // min, NaN-propagating like numpy.min
        for (size_t k = BUMP_OUT_HEADER; k < out.size(); ++k)
            if (std::isnan(out[k])) neff_min = NAN;
        rec[0] = loglike, rec[1] = -nobs * log_mu, rec[2] = out[BUMP_OUT_NEFF_SEL];
        rec[3] = nobs / mu + sqrt(nobs) / mu * x[14];   // R (:396-399)
        rec[4] = theta[7], rec[5] = fpl, rec[6] = theta[12], rec[7] = neff_min;
        if (!(std::isfinite(logl) && std::isfinite(lpj))) {
            for (int i = 0; i < NSITES; ++i) g[i] = 0.0;
            return INF;
        }
        double gt[BUMP_NTHETA];
        for (int k = 0; k < BUMP_NTHETA; ++k) gt[k] = out[BUMP_OUT_DLOGLIKE + k] - nobs * out[BUMP_OUT_DLOG_MU + k];
        // chain rule to the sites (priors.grad_sites_from_theta)
        const double gs[NSITES] = {gt[0], gt[1], gt[2], gt[3], gt[4], gt[5], gt[6] + gt[7], gt[7], gt[8], gt[10],
                                   fpl * gt[9], gt[11] + gt[12], gt[12], gt[13], 0.0};
        for (int i = 0; i < NSITES; ++i) g[i] = -((gs[i] + glp[i]) * dx[i] + dlj[i]);
        return -(lpj + logl);
    }
};

// ------------------------------------------------------------------ adaptation (nuts.py DualAveraging, Welford, windows)
struct DualAveraging {
    double mu, target, gamma = 0.05, t0 = 10.0, kappa = 0.75, hbar = 0.0, log_eps, log_eps_bar = 0.0;
    int t = 0;
    DualAveraging(const double eps0, const double target_) : mu(log(10.0 * eps0)), target(target_), log_eps(log(eps0)) {}
    double update(const double accept) {
        ++t;
        const double eta = 1.0 / (t + t0);
        hbar = (1 - eta) * hbar + eta * (target - accept);
        log_eps = mu - sqrt((double)t) / gamma * hbar;
        const double w = pow((double)t, -kappa);
        log_eps_bar = w * log_eps + (1 - w) * log_eps_bar;
        return exp(log_eps);
    }
    double final_eps() const { return exp(log_eps_bar); }
};

struct Welford {
    int d, n = 0;
    std::vector<double> mean, m2;
    explicit Welford(const int d_) : d(d_), mean(d_, 0.0), m2((size_t)d_ * d_, 0.0) {}
    void add(const double* x) {
        ++n;
        double dl[MAXD], dr[MAXD];
        for (int i = 0; i < d; ++i) {
            dl[i] = x[i] - mean[i];
            mean[i] += dl[i] / n;
            dr[i] = x[i] - mean[i];
        }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) m2[(size_t)i * d + j] += dl[i] * dr[j];
    }
    void covariance(double* cov) const {   // with Stan's shrinkage towards 1e-3 I
        const double a = n / (n + 5.0), b = 1e-3 * (5.0 / (n + 5.0));
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                cov[i * MAXD + j] = a * m2[(size_t)i * d + j] / std::max(n - 1, 1) + (i == j ? b : 0.0);
    }
};

void adaptation_windows(const int num_warmup, int& slow_start, std::vector<int>& ends) {
    int init_buffer = 75, term_buffer = 50, base_window = 25;
    ends.clear();
    slow_start = 0;
    if (num_warmup < 20) return;
    if (init_buffer + base_window + term_buffer > num_warmup) {
        init_buffer = (int)(0.15 * num_warmup);
        term_buffer = (int)(0.10 * num_warmup);
        base_window = num_warmup - init_buffer - term_buffer;
    }
    int start = init_buffer, size = base_window;
    const int last = num_warmup - term_buffer;
    while (start < last) {
        int end = start + size;
        if (end + 2 * size > last) end = last;
        ends.push_back(end);
        start = end;
        size *= 2;
    }
    slow_start = init_buffer;
}

// ------------------------------------------------------------------ the sampler (nuts.py NUTS)
struct State {
    double u[MAXD], p[MAXD], g[MAXD], U;
    double rec[NREC];
};
struct Tree {
    State left, right, prop;
    double lw, rho[MAXD], sum_acc;
    bool turning, diverging;
    int n;
};

double logaddexp(const double a, const double b) {
    if (a == -INF) return b;
    if (b == -INF) return a;
    const double m = std::max(a, b);
    return m + log1p(exp(-fabs(a - b)));
}

struct Sampler {
    Potential& f;
    const int d;
    const bool dense;
    const int max_depth;
    double minv[MAXD * MAXD], chol[MAXD * MAXD];   // inverse mass ~ posterior covariance = chol chol^T; p ~ N(0, M)
    std::mt19937_64 rng;
    std::normal_distribution<double> normal{0.0, 1.0};
    std::uniform_real_distribution<double> unif{0.0, 1.0};
    long n_leapfrog = 0;

    Sampler(Potential& f_, const bool dense_, const int max_depth_, const uint64_t seed)
        : f(f_), d(f_.dim), dense(dense_), max_depth(max_depth_), rng(seed) {
        double eye[MAXD * MAXD] = {};
        for (int i = 0; i < d; ++i) eye[i * MAXD + i] = 1.0;
        set_mass(eye);
    }
    bool set_mass(const double* m) {
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) minv[i * MAXD + j] = (dense || i == j) ? m[i * MAXD + j] : 0.0;
        for (int i = 0; i < d; ++i) {   // Cholesky, lower
            for (int j = 0; j <= i; ++j) {
                double s = minv[i * MAXD + j];
                for (int k = 0; k < j; ++k) s -= chol[i * MAXD + k] * chol[j * MAXD + k];
                if (i == j) {
                    if (!(s > 0.0)) return false;
                    chol[i * MAXD + i] = sqrt(s);
                } else {
                    chol[i * MAXD + j] = s / chol[j * MAXD + j];
                }
            }
            for (int j = i + 1; j < d; ++j) chol[i * MAXD + j] = 0.0;
        }
        return true;
    }
    void draw_momentum(double* p) {   // solve chol^T p = z
        double z[MAXD];
        for (int i = 0; i < d; ++i) z[i] = normal(rng);
        for (int i = d - 1; i >= 0; --i) {
            double s = z[i];
            for (int k = i + 1; k < d; ++k) s -= chol[k * MAXD + i] * p[k];
            p[i] = s / chol[i * MAXD + i];
        }
    }
    void minv_mul(const double* p, double* o) const {
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += minv[i * MAXD + j] * p[j];
            o[i] = s;
        }
    }
    double dot(const double* a, const double* b) const {
        double s = 0.0;
        for (int i = 0; i < d; ++i) s += a[i] * b[i];
        return s;
    }
    double kinetic(const double* p) const {
        double v[MAXD];
        minv_mul(p, v);
        return 0.5 * dot(p, v);
    }
    void leapfrog(const State& s, const double eps, State& o) {
        double v[MAXD];
        for (int i = 0; i < d; ++i) o.p[i] = s.p[i] - 0.5 * eps * s.g[i];
        minv_mul(o.p, v);
        for (int i = 0; i < d; ++i) o.u[i] = s.u[i] + eps * v[i];
        o.U = f.eval(o.u, o.g, o.rec);
        ++n_leapfrog;
        if (std::isfinite(o.U))
            for (int i = 0; i < d; ++i) o.p[i] -= 0.5 * eps * o.g[i];
    }
    double energy(const State& s) const { return std::isfinite(s.U) ? s.U + kinetic(s.p) : INF; }

    // Hoffman & Gelman alg. 4: double / halve until the one-step acceptance crosses 0.8
    double find_reasonable_step_size(const State& at, double eps) {
        auto accept_logp = [&](const double e) {
            State s0 = at, s1;
            draw_momentum(s0.p);
            leapfrog(s0, e, s1);
            return (at.U + kinetic(s0.p)) - energy(s1);
        };
        const double target = log(0.8);
        double dlt = accept_logp(eps);
        const int direction = dlt > target ? 1 : -1;
        for (int it = 0; it < 50; ++it) {
            eps *= direction == 1 ? 2.0 : 0.5;
            dlt = accept_logp(eps);
            if ((direction == 1 && !(dlt > target)) || (direction == -1 && dlt > target)) break;
        }
        return eps;
    }
    bool uturn(const double* rho, const double* pl, const double* pr) const {
        double v[MAXD];
        minv_mul(pl, v);
        if (dot(rho, v) <= 0) return true;
        minv_mul(pr, v);
        return dot(rho, v) <= 0;
    }
    bool uturn_sum(const double* a, const double* b, const double* pl, const double* pr) const {   // rho = a + b
        double rho[MAXD];
        for (int i = 0; i < d; ++i) rho[i] = a[i] + b[i];
        return uturn(rho, pl, pr);
    }
    // recursive doubling, multinomial sampling, generalised U-turn with the extra junction checks (Stan >= 2.20)
    void build(const State& s, const int direction, const int depth, const double eps, const double h0, Tree& t) {
        if (depth == 0) {
            leapfrog(s, direction * eps, t.left);
            double dh = energy(t.left) - h0;
            if (std::isnan(dh)) dh = INF;
            t.right = t.left;
            t.prop = t.left;
            t.lw = -dh;
            memcpy(t.rho, t.left.p, sizeof(double) * d);
            t.turning = false;
            t.diverging = dh > MAX_DELTA_H;
            t.sum_acc = dh > -700 ? std::min(1.0, exp(-dh)) : 1.0;
            t.n = 1;
            return;
        }
        build(s, direction, depth - 1, eps, h0, t);
        if (t.turning || t.diverging) return;
        Tree* t2 = new Tree;   // off the stack: ten levels of ~2.5 KB trees would still be fine, this keeps it flat
        build(direction == 1 ? t.right : t.left, direction, depth - 1, eps, h0, *t2);
        const double lw = logaddexp(t.lw, t2->lw);
        // l1, r1 = t.left, t.right;  l2, r2 = t2->left, t2->right
        bool take2 = false;
        if (!(t2->turning || t2->diverging)) take2 = log(unif(rng)) < t2->lw - lw;
        bool turning;
        {
            double rho[MAXD];
            for (int i = 0; i < d; ++i) rho[i] = t.rho[i] + t2->rho[i];
            const State& left = direction == 1 ? t.left : t2->left;
            const State& right = direction == 1 ? t2->right : t.right;
            turning = t2->turning || uturn(rho, left.p, right.p);
            if (!turning) {
                if (direction == 1)
                    turning = uturn_sum(t.rho, t2->left.p, t.left.p, t2->left.p) ||
                              uturn_sum(t2->rho, t.right.p, t.right.p, t2->right.p);
                else
                    turning = uturn_sum(t2->rho, t.left.p, t2->left.p, t.left.p) ||
                              uturn_sum(t.rho, t2->right.p, t2->right.p, t.right.p);
            }
            memcpy(t.rho, rho, sizeof(double) * d);
        }
        if (take2) t.prop = t2->prop;
        if (direction == 1) t.right = t2->right;
        else t.left = t2->left;
        t.lw = lw;
        t.turning = turning;
        t.diverging = t2->diverging;
        t.sum_acc += t2->sum_acc;
        t.n += t2->n;
        delete t2;
    }
    // one transition from `cur` (u, U, g, rec); returns the new state in `cur`
    void transition(State& cur, const double eps, double& accept, bool& diverging, int& depth, int& n) {
        draw_momentum(cur.p);
        const double h0 = cur.U + kinetic(cur.p);
        State left = cur, right = cur, prop = cur;
        double lw = 0.0, rho[MAXD], sum_acc = 0.0;
        memcpy(rho, cur.p, sizeof(double) * d);
        n = 0, depth = 0, diverging = false;
        Tree* t2 = new Tree;
        while (depth < max_depth) {
            const int direction = unif(rng) < 0.5 ? 1 : -1;
            build(direction == 1 ? right : left, direction, depth, eps, h0, *t2);
            sum_acc += t2->sum_acc;
            n += t2->n;
            if (t2->diverging) {
                diverging = true;
                break;
            }
            if (t2->turning) break;
            if (log(unif(rng)) < t2->lw - lw) prop = t2->prop;   // biased progressive sampling at the top level
            bool extra;
            if (direction == 1) {
                extra = uturn_sum(rho, t2->left.p, left.p, t2->left.p) || uturn_sum(t2->rho, right.p, right.p, t2->right.p);
                right = t2->right;
            } else {
                extra = uturn_sum(t2->rho, left.p, t2->left.p, left.p) || uturn_sum(rho, t2->right.p, t2->right.p, right.p);
                left = t2->left;
            }
            lw = logaddexp(lw, t2->lw);
            for (int i = 0; i < d; ++i) rho[i] += t2->rho[i];
            ++depth;
            if (extra || uturn(rho, left.p, right.p)) break;
        }
        delete t2;
        cur = prop;
        accept = sum_acc / std::max(n, 1);
    }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// nuts.py run_chain
int run_chain(Potential& f, const int num_warmup, const int num_samples, const uint64_t seed, const int dense_mass,
              const double target_accept, const int max_tree_depth, const double* init_u, double* out_u, double* out_x,
              double* out_stats, double* out_det, double* out_info, double* out_minv) {
    const int d = f.dim;
    if (d < 1 || d > MAXD) return bump_set_error(BUMP_E_INVALID, "NUTS: dimension must be in 1..32");
    if (num_warmup < 0 || num_samples < 0 || !out_u || !out_stats || !out_info)
        return bump_set_error(BUMP_E_INVALID, "NUTS: bad arguments");
    Sampler k(f, dense_mass != 0, max_tree_depth, seed);
    State cur;
    memset(&cur, 0, sizeof(cur));
    // init like numpyro's init_to_uniform: uniform(-2, 2) in unconstrained space, retried until finite
    bool found = false;
    for (int attempt = 0; attempt < 100 && !found; ++attempt) {
        for (int i = 0; i < d; ++i) cur.u[i] = (init_u && attempt == 0) ? init_u[i] : -2.0 + 4.0 * k.unif(k.rng);
        cur.U = f.eval(cur.u, cur.g, cur.rec);
        found = std::isfinite(cur.U);
        if (f.error()) return f.error();
    }
    if (!found) return bump_set_error(BUMP_E_INVALID, "NUTS: could not find a finite starting point");
    double eps = k.find_reasonable_step_size(cur, 1.0);
    DualAveraging da(eps, target_accept);
    int slow_start = 0;
    std::vector<int> windows;
    adaptation_windows(num_warmup, slow_start, windows);
    Welford wf(d);
    const int total = num_warmup + num_samples;
    const double t_start = now_s();
    double t_warm = -1.0;
    for (int it = 0; it < total; ++it) {
        if (it == num_warmup) {
            eps = da.final_eps();
            t_warm = now_s();
        }
        double acc;
        bool div;
        int depth, nlf;
        k.transition(cur, eps, acc, div, depth, nlf);
        if (f.error()) return f.error();
        if (it < num_warmup) {
            eps = da.update(acc);
            if (!windows.empty() && slow_start <= it && it < windows.back()) wf.add(cur.u);
            bool window_end = false;
            for (int e : windows) window_end |= (it + 1 == e);
            if (window_end) {
                double cov[MAXD * MAXD];
                wf.covariance(cov);
                if (!k.set_mass(cov)) return bump_set_error(BUMP_E_INVALID, "NUTS: adapted mass matrix is not positive definite");
                wf = Welford(d);
                eps = k.find_reasonable_step_size(cur, da.final_eps());
                da = DualAveraging(eps, target_accept);
            }
        } else {
            const int j = it - num_warmup;
            memcpy(out_u + (size_t)j * d, cur.u, sizeof(double) * d);
            if (out_x) f.constrain(cur.u, out_x + (size_t)j * d);
            double* st = out_stats + (size_t)j * BUMP_NUTS_NSTAT;
            st[0] = acc, st[1] = depth, st[2] = nlf, st[3] = div ? 1.0 : 0.0, st[4] = cur.U;
            if (out_det) memcpy(out_det + (size_t)j * NREC, cur.rec, sizeof(double) * NREC);
        }
    }
    const double t_end = now_s();
    if (t_warm < 0) t_warm = t_end;
    out_info[0] = eps;
    out_info[1] = (double)k.n_leapfrog;
    out_info[2] = t_warm - t_start;
    out_info[3] = t_end - t_warm;
    out_info[4] = (double)f.n_evals;
    if (out_minv)
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) out_minv[i * d + j] = k.minv[i * MAXD + j];
    return BUMP_OK;
}

}  // namespace

extern "C" {

int bump_nuts_chain_cb(bump_potential_cb f, void* user, int dim, int num_warmup, int num_samples, uint64_t seed,
                       int dense_mass, double target_accept, int max_tree_depth, const double* init_u, double* out_u,
                       double* out_stats, double* out_info, double* out_minv) {
    if (!f) return bump_set_error(BUMP_E_INVALID, "NUTS: null potential");
    CallbackPotential p(f, user, dim);
    return run_chain(p, num_warmup, num_samples, seed, dense_mass, target_accept, max_tree_depth, init_u, out_u, nullptr,
                     out_stats, nullptr, out_info, out_minv);
}

int bump_nuts_prior_terms(const double* u, double* x, double* prior_u, double* prior_grad) {
    if (!u || !x || !prior_u || !prior_grad) return bump_set_error(BUMP_E_INVALID, "NUTS: null argument");
    double total = 0.0;
    for (int i = 0; i < NSITES; ++i) {
        double dx, dlj, glp;
        total += site_terms(sites()[i], u[i], x[i], dx, dlj, glp);
        prior_grad[i] = -(glp * dx + dlj);
    }
    *prior_u = -total;
    return BUMP_OK;
}

int bump_nuts_potential(bump_ctx* ctx, const double* u, double* U, double* grad, double* rec) {
    if (!ctx || !u || !U || !grad) return bump_set_error(BUMP_E_INVALID, "NUTS: null argument");
    if (bump_ctx_flags(ctx) & (BUMP_FLAG_WA | BUMP_FLAG_FIXED_COSMO))
        return bump_set_error(BUMP_E_INVALID, "NUTS: the driver binds pop_cosmo_model (no w0-wa / fixed-cosmology contexts)");
    ModelPotential p(ctx);
    double r[NREC];
    *U = p.eval(u, grad, r);
    if (rec) memcpy(rec, r, sizeof(r));
    return p.error();
}

int bump_nuts_chain(bump_ctx* ctx, int num_warmup, int num_samples, uint64_t seed, int dense_mass, double target_accept,
                    int max_tree_depth, const double* init_u, double* out_u, double* out_x, double* out_stats,
                    double* out_det, double* out_info, double* out_minv) {
    if (!ctx) return bump_set_error(BUMP_E_INVALID, "NUTS: null context");
    if (bump_ctx_flags(ctx) & (BUMP_FLAG_WA | BUMP_FLAG_FIXED_COSMO))
        return bump_set_error(BUMP_E_INVALID, "NUTS: the driver binds pop_cosmo_model (no w0-wa / fixed-cosmology contexts)");
    ModelPotential p(ctx);
    return run_chain(p, num_warmup, num_samples, seed, dense_mass, target_accept, max_tree_depth, init_u, out_u, out_x,
                     out_stats, out_det, out_info, out_minv);
}

}  // extern "C"

"""Host NUTS driver for the bound model.

The reference samples with numpyro: `NUTS(pop_cosmo_model, dense_mass=True)`, `MCMC(num_warmup=1000,
num_samples=1000, num_chains=4)`, seed 1652819403 (/root/reference/src/scripts/run_cosmo_fit.py:17-19,45-49).
numpyro is not installable in this image, so this module restates the algorithm it runs (third-party, unpinned):
multinomial NUTS with the generalised U-turn criterion, max tree depth 10, divergence threshold 1000, dual-averaging
step-size adaptation to a target acceptance of 0.8, and Stan-style windowed adaptation of a dense mass matrix.
Every leapfrog step is one call of `model.potential(u)`, i.e. one evaluation of the CUDA hot path.

Not bit-compatible with numpyro's random stream (JAX threefry vs numpy PCG64); the acceptance test is statistical
(posterior moments and ESS), see tests/test_nuts.py.
"""
import math
import time

import numpy as np

MAX_DELTA_H = 1000.0


class DualAveraging:
    """Nesterov dual averaging of log step size (Hoffman & Gelman 2014, numpyro/Stan defaults)."""

    def __init__(self, eps0, target=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.mu = math.log(10.0 * eps0)
        self.target, self.gamma, self.t0, self.kappa = target, gamma, t0, kappa
        self.t, self.hbar, self.log_eps, self.log_eps_bar = 0, 0.0, math.log(eps0), 0.0

    def update(self, accept_prob):
        self.t += 1
        eta = 1.0 / (self.t + self.t0)
        self.hbar = (1 - eta) * self.hbar + eta * (self.target - accept_prob)
        self.log_eps = self.mu - math.sqrt(self.t) / self.gamma * self.hbar
        w = self.t ** (-self.kappa)
        self.log_eps_bar = w * self.log_eps + (1 - w) * self.log_eps_bar
        return math.exp(self.log_eps)

    def final(self):
        return math.exp(self.log_eps_bar)


class Welford:
    def __init__(self, d):
        self.n, self.mean, self.m2 = 0, np.zeros(d), np.zeros((d, d))

    def add(self, x):
        self.n += 1
        d = x - self.mean
        self.mean += d / self.n
        self.m2 += np.outer(d, x - self.mean)

    def covariance(self, regularize=True):
        cov = self.m2 / max(self.n - 1, 1)
        if regularize:   # Stan's shrinkage towards 1e-3 * I
            cov = (self.n / (self.n + 5.0)) * cov + 1e-3 * (5.0 / (self.n + 5.0)) * np.eye(cov.shape[0])
        return cov


def adaptation_windows(num_warmup, init_buffer=75, term_buffer=50, base_window=25):
    """(start of the first slow window, end indices (exclusive) of the slow (mass-matrix) windows):
    Stan/numpyro schedule."""
    if num_warmup < 20:
        return 0, []
    if init_buffer + base_window + term_buffer > num_warmup:
        init_buffer = int(0.15 * num_warmup)
        term_buffer = int(0.10 * num_warmup)
        base_window = num_warmup - init_buffer - term_buffer
    ends, start, size = [], init_buffer, base_window
    last = num_warmup - term_buffer
    while start < last:
        end = start + size
        if end + 2 * size > last:
            end = last
        ends.append(end)
        start, size = end, 2 * size
    return init_buffer, ends


class _State:
    __slots__ = ("u", "p", "U", "g", "ev")

    def __init__(self, u, p, U, g, ev):
        self.u, self.p, self.U, self.g, self.ev = u, p, U, g, ev


class NUTS:
    def __init__(self, potential_fn, dim, dense_mass=True, target_accept=0.8, max_tree_depth=10, rng=None):
        self.f, self.dim, self.dense = potential_fn, dim, dense_mass
        self.target, self.max_depth = target_accept, max_tree_depth
        self.rng = rng if rng is not None else np.random.default_rng()
        self.set_mass(np.eye(dim))
        self.n_leapfrog = 0

    # inverse mass matrix Minv ~ posterior covariance; p ~ N(0, M)
    def set_mass(self, minv):
        self.minv = minv if self.dense else np.diag(np.diag(minv))
        self.chol_minv = np.linalg.cholesky(self.minv)

    def draw_momentum(self):
        z = self.rng.standard_normal(self.dim)
        return np.linalg.solve(self.chol_minv.T, z)

    def kinetic(self, p):
        return 0.5 * float(p @ (self.minv @ p))

    def leapfrog(self, s, eps):
        p = s.p - 0.5 * eps * s.g
        u = s.u + eps * (self.minv @ p)
        U, g, ev = self.f(u)
        self.n_leapfrog += 1
        if math.isfinite(U):
            p = p - 0.5 * eps * g
        return _State(u, p, U, g, ev)

    def find_reasonable_step_size(self, u, U, g, eps=1.0):
        """Heuristic of Hoffman & Gelman (alg. 4): double/halve until the one-step acceptance crosses 0.8."""
        def accept_logp(e):
            p0 = self.draw_momentum()
            s = self.leapfrog(_State(u, p0, U, g, None), e)
            h0 = U + self.kinetic(p0)
            h1 = s.U + self.kinetic(s.p) if math.isfinite(s.U) else math.inf
            return h0 - h1
        target = math.log(0.8)
        d = accept_logp(eps)
        direction = 1 if d > target else -1
        for _ in range(50):
            eps = eps * (2.0 ** direction)
            d = accept_logp(eps)
            if (direction == 1 and not d > target) or (direction == -1 and d > target):
                break
        return eps

    # ---- one NUTS transition (recursive doubling, multinomial sampling, generalised U-turn)
    def _uturn(self, rho, p_left, p_right):
        return (rho @ (self.minv @ p_left) <= 0) or (rho @ (self.minv @ p_right) <= 0)

    def _build(self, s, direction, depth, eps, h0):
        """Returns (left, right, proposal, log_w, rho, turning, diverging, sum_accept, n)."""
        if depth == 0:
            s1 = self.leapfrog(s, direction * eps)
            h1 = s1.U + self.kinetic(s1.p) if math.isfinite(s1.U) else math.inf
            dh = h1 - h0
            if math.isnan(dh):
                dh = math.inf
            diverging = dh > MAX_DELTA_H
            acc = min(1.0, math.exp(-dh)) if dh > -700 else 1.0
            return s1, s1, s1, -dh, s1.p.copy(), False, diverging, acc, 1
        l1, r1, prop1, lw1, rho1, turn1, div1, acc1, n1 = self._build(s, direction, depth - 1, eps, h0)
        if turn1 or div1:
            return l1, r1, prop1, lw1, rho1, turn1, div1, acc1, n1
        edge = r1 if direction == 1 else l1
        l2, r2, prop2, lw2, rho2, turn2, div2, acc2, n2 = self._build(edge, direction, depth - 1, eps, h0)
        left, right = (l1, r2) if direction == 1 else (l2, r1)
        lw = np.logaddexp(lw1, lw2)
        prop = prop2 if (not (turn2 or div2)) and math.log(self.rng.uniform()) < lw2 - lw else prop1
        rho = rho1 + rho2
        turning = turn2 or self._uturn(rho, left.p, right.p)
        if not turning:   # extra checks across the junction of the two subtrees (Stan >= 2.20)
            if direction == 1:
                turning = self._uturn(rho1 + l2.p, l1.p, l2.p) or self._uturn(rho2 + r1.p, r1.p, r2.p)
            else:
                turning = self._uturn(rho2 + l1.p, l2.p, l1.p) or self._uturn(rho1 + r2.p, r2.p, r1.p)
        return left, right, prop, lw, rho, turning, div2, acc1 + acc2, n1 + n2

    def transition(self, u, U, g, ev, eps):
        p0 = self.draw_momentum()
        h0 = U + self.kinetic(p0)
        s0 = _State(u, p0, U, g, ev)
        left = right = prop = s0
        lw, rho = 0.0, p0.copy()
        sum_acc, n, diverging, depth = 0.0, 0, False, 0
        while depth < self.max_depth:
            direction = 1 if self.rng.uniform() < 0.5 else -1
            edge = right if direction == 1 else left
            l2, r2, prop2, lw2, rho2, turn2, div2, acc2, n2 = self._build(edge, direction, depth, eps, h0)
            sum_acc += acc2
            n += n2
            if div2:
                diverging = True
                break
            if turn2:
                break
            if math.log(self.rng.uniform()) < lw2 - lw:   # biased progressive sampling at the top level
                prop = prop2
            if direction == 1:
                extra = self._uturn(rho + l2.p, left.p, l2.p) or self._uturn(rho2 + right.p, right.p, r2.p)
                right = r2
            else:
                extra = self._uturn(rho2 + left.p, l2.p, left.p) or self._uturn(rho + r2.p, r2.p, right.p)
                left = l2
            lw = np.logaddexp(lw, lw2)
            rho = rho + rho2
            depth += 1
            if extra or self._uturn(rho, left.p, right.p):
                break
        return prop, sum_acc / max(n, 1), diverging, depth, n


def run_chain(model, num_warmup=1000, num_samples=1000, seed=0, dense_mass=True, target_accept=0.8,
              max_tree_depth=10, init=None, progress=None):
    """One chain.  `model.potential(u)` -> (U, dU/du, evaluation dict).  Returns a dict of arrays."""
    from . import priors
    rng = np.random.default_rng(seed)
    dim = getattr(model, "dim", priors.NSITES)          # pop_cosmo_model: 15 sites; pop_model: 12
    constrain = getattr(model, "constrain", priors.constrain)
    f = model.potential
    # init like numpyro's init_to_uniform: uniform(-2, 2) in unconstrained space, retried until finite
    for _ in range(100):
        u = rng.uniform(-2, 2, dim) if init is None else np.asarray(init, dtype=np.float64)
        U, g, ev = f(u)
        if math.isfinite(U) and np.all(np.isfinite(g)):
            break
        init = None
    else:
        raise RuntimeError("could not find a finite starting point")
    k = NUTS(f, dim, dense_mass, target_accept, max_tree_depth, rng)
    eps = k.find_reasonable_step_size(u, U, g)
    da = DualAveraging(eps, target_accept)
    slow_start, windows = adaptation_windows(num_warmup)
    wf = Welford(dim)
    total = num_warmup + num_samples
    out_u = np.empty((num_samples, dim))
    out_x = np.empty((num_samples, dim))
    stats = {k_: np.empty(num_samples) for k_ in ("accept", "depth", "n_leapfrog", "diverging", "potential")}
    det = {k_: [] for k_ in ("loglike", "selfactor", "neff_sel", "R", "mbhmax", "fpl", "kappa", "neff_min")}
    t_start = time.perf_counter()
    t_warm = None
    for it in range(total):
        if it == num_warmup:
            eps = da.final()
            t_warm = time.perf_counter()
        s, acc, div, depth, nlf = k.transition(u, U, g, ev, eps)
        u, U, g, ev = s.u, s.U, s.g, s.ev
        if it < num_warmup:
            eps = da.update(acc)
            if windows and slow_start <= it < windows[-1]:
                wf.add(u)
            if it + 1 in windows:
                k.set_mass(wf.covariance())
                wf = Welford(dim)
                eps = k.find_reasonable_step_size(u, U, g, da.final())
                da = DualAveraging(eps, target_accept)
        else:
            j = it - num_warmup
            out_u[j] = u
            out_x[j] = constrain(u)[0]
            stats["accept"][j], stats["depth"][j], stats["n_leapfrog"][j] = acc, depth, nlf
            stats["diverging"][j], stats["potential"][j] = div, U
            evd = model.deterministics(ev) if hasattr(model, "deterministics") else ev
            for name in ("loglike", "selfactor", "neff_sel", "R", "mbhmax", "fpl", "kappa"):
                det[name].append(evd[name])
            det["neff_min"].append(float(np.min(evd["neff"])) if len(evd["neff"]) else float("nan"))
        if progress and (it + 1) % progress == 0:
            print(f"  chain seed {seed}: {it + 1}/{total} eps={eps:.4f} depth={depth} acc={acc:.2f}", flush=True)
    t_end = time.perf_counter()
    return {"u": out_u, "x": out_x, "stats": stats, "deterministic": {a: np.array(b) for a, b in det.items()},
            "step_size": eps, "inverse_mass": k.minv, "n_leapfrog_total": k.n_leapfrog,
            "warmup_s": (t_warm or t_end) - t_start, "sampling_s": t_end - (t_warm or t_end)}


# ------------------------------------------------------------------ the same driver in C++ (csrc/bump_nuts.cpp)
_STAT_NAMES = ("accept", "depth", "n_leapfrog", "diverging", "potential")
_DET_NAMES = ("loglike", "selfactor", "neff_sel", "R", "mbhmax", "fpl", "kappa", "neff_min")


def _native_result(dim, num_samples, u, x, stats, det, info, minv):
    return {"u": u, "x": x, "stats": {k: stats[:, i].copy() for i, k in enumerate(_STAT_NAMES)},
            "deterministic": {k: det[:, i].copy() for i, k in enumerate(_DET_NAMES)},
            "step_size": float(info[0]), "inverse_mass": minv, "n_leapfrog_total": int(info[1]),
            "warmup_s": float(info[2]), "sampling_s": float(info[3]), "n_evals": int(info[4])}


def run_chain_native(model, num_warmup=1000, num_samples=1000, seed=0, dense_mass=True, target_accept=0.8,
                     max_tree_depth=10, init=None, progress=None):
    """One chain run by the library's C++ driver (`bump_nuts_chain`): same algorithm as `run_chain`, no interpreter in
    the leapfrog loop.  `model` is a bound `pop_cosmo_model` (single-rank context).  The call releases the GIL, so
    chains started from several Python threads run truly in parallel."""
    from . import _lib, priors
    ctx = getattr(getattr(model, "like", None), "_ctx", None)   # ShardedHyperlikelihood has none: every rank would
    # have to take the same steps in lock step, which the Python driver does through the merged potential
    if ctx is None:
        raise TypeError("run_chain_native needs a model bound to a single-rank Hyperlikelihood")
    lib = _lib.load()
    d = priors.NSITES
    u = np.empty((num_samples, d))
    x = np.empty((num_samples, d))
    stats = np.empty((num_samples, _lib.NUTS_NSTAT))
    det = np.empty((num_samples, _lib.NUTS_NDET))
    info = np.zeros(8)
    minv = np.empty((d, d))
    init_p = _lib.as_dp(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else None
    _lib.check(lib.bump_nuts_chain(ctx, num_warmup, num_samples, int(seed) & (2**64 - 1), int(dense_mass),
                                   float(target_accept), int(max_tree_depth), init_p, _lib.as_dp(u), _lib.as_dp(x),
                                   _lib.as_dp(stats), _lib.as_dp(det), _lib.as_dp(info), _lib.as_dp(minv)))
    out = _native_result(d, num_samples, u, x, stats, det, info, minv)
    if hasattr(model, "n_evals"):
        model.n_evals += out["n_evals"]
    return out


def run_chain_native_fn(potential_fn, dim, num_warmup=1000, num_samples=1000, seed=0, dense_mass=True,
                        target_accept=0.8, max_tree_depth=10, init=None):
    """The C++ driver on an arbitrary Python potential `u -> (U, grad)` (`bump_nuts_chain_cb`; for tests: the
    callback re-enters the interpreter on every leapfrog step)."""
    from . import _lib
    lib = _lib.load()

    def cb(_user, u_p, g_p):
        uu = np.ctypeslib.as_array(u_p, shape=(dim,))
        U, g = potential_fn(uu.copy())[:2]
        np.ctypeslib.as_array(g_p, shape=(dim,))[:] = g
        return float(U)

    cfn = _lib.POTENTIAL_CB(cb)
    u = np.empty((num_samples, dim))
    stats = np.empty((num_samples, _lib.NUTS_NSTAT))
    info = np.zeros(8)
    minv = np.empty((dim, dim))
    init_p = _lib.as_dp(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else None
    import ctypes
    _lib.check(lib.bump_nuts_chain_cb(ctypes.cast(cfn, ctypes.c_void_p), None, dim, num_warmup, num_samples,
                                      int(seed) & (2**64 - 1), int(dense_mass), float(target_accept),
                                      int(max_tree_depth), init_p, _lib.as_dp(u), _lib.as_dp(stats), _lib.as_dp(info),
                                      _lib.as_dp(minv)))
    return _native_result(dim, num_samples, u, u.copy(), stats, np.zeros((num_samples, _lib.NUTS_NDET)), info, minv)


def run_mcmc(model, num_warmup=1000, num_samples=1000, num_chains=4, seed=1652819403, native=False, **kw):
    """The reference's MCMC configuration (run_cosmo_fit.py:17-19,45-46).

    `model` is one bound model (chains run one after another on it) or a list of `num_chains` bound models, one per
    chain, which then run concurrently in threads — the reference's chains are parallel too (numpyro pmaps them over
    host devices, run_cosmo_fit.py:1-3).  The C call releases the GIL, so one chain's host work (priors,
    transforms, tree bookkeeping) overlaps the others' GPU evaluations; evaluations of different contexts on one
    device overlap as long as they sit on different constant-bank slots (four per device).  `native=True` runs each
    chain in the library's C++ driver (`run_chain_native`) instead of the Python one."""
    t0 = time.perf_counter()
    chain_fn = run_chain_native if native else run_chain
    if isinstance(model, (list, tuple)):
        import threading
        models = list(model)
        if len(models) != num_chains:
            raise ValueError("need one model per chain")
        chains = [None] * num_chains
        errors = []

        def work(c):
            try:
                chains[c] = chain_fn(models[c], num_warmup, num_samples, seed=seed + c, **kw)
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=work, args=(c,)) for c in range(num_chains)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        warm = max(c["warmup_s"] for c in chains)
        samp = max(c["sampling_s"] for c in chains)
    else:
        chains = [chain_fn(model, num_warmup, num_samples, seed=seed + c, **kw) for c in range(num_chains)]
        warm = sum(c["warmup_s"] for c in chains)
        samp = sum(c["sampling_s"] for c in chains)
    x = np.stack([c["x"] for c in chains])                      # [chain, draw, site]
    return {"x": x, "chains": chains, "ess_bulk": np.array([ess_bulk(x[:, :, i]) for i in range(x.shape[2])]),
            "rhat": np.array([split_rhat(x[:, :, i]) for i in range(x.shape[2])]),
            "warmup_s": warm, "sampling_s": samp, "wall_s": time.perf_counter() - t0,
            "n_leapfrog_total": sum(c["n_leapfrog_total"] for c in chains)}


# ------------------------------------------------------------------ diagnostics (arviz-equivalent formulas)
def _split(x):
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, n:2 * n]], axis=0)


def _rank_normalize(x):
    from scipy import stats
    r = stats.rankdata(x.ravel(), method="average").reshape(x.shape)
    return stats.norm.ppf((r - 0.375) / (x.size + 0.25))


def _autocov(x):
    n = x.shape[-1]
    m = 1 << (2 * n - 1).bit_length()
    f = np.fft.rfft(x - x.mean(axis=-1, keepdims=True), m, axis=-1)
    return np.fft.irfft(f * np.conj(f), m, axis=-1)[..., :n] / n


def _ess(x):
    """ESS of [chains, draws] with Geyer's initial monotone sequence (Vehtari et al. 2021)."""
    m, n = x.shape
    acov = _autocov(x)
    chain_var = acov[:, 0] * n / (n - 1.0)
    mean_var = chain_var.mean()
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    rho = np.zeros(n)
    t = 0
    rho_even, rho[0] = 1.0, 1.0
    rho_odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n - 3 and rho_even + rho_odd > 0:
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1], rho[t + 2] = rho_even, rho_odd
        t += 2
    max_t = t
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:   # monotone
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + (rho[max_t + 1] if rho_even > 0 else 0.0)
    tau = max(tau, 1.0 / math.log10(m * n))
    return m * n / tau


def ess_bulk(x):
    return float(_ess(_rank_normalize(_split(np.asarray(x, dtype=np.float64)))))


def split_rhat(x):
    z = _rank_normalize(_split(np.asarray(x, dtype=np.float64)))
    n = z.shape[1]
    w = z.var(axis=1, ddof=1).mean()
    b = n * z.mean(axis=1).var(ddof=1)
    return float(math.sqrt(((n - 1.0) / n * w + b / n) / w))

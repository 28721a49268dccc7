"""Priors and constrained <-> unconstrained transforms of the reference's sample sites, on the host.

Mirrors /root/reference/src/scripts/intensity_models.py:281-311 (mass_parameters, redshift_parameters,
cosmo_parameters) and :398 (R_unit).  numpyro semantics restated (third-party, unpinned by the reference):
TruncatedNormal log-density with the Phi normaliser; NUTS works in unconstrained space through
`biject_to(support)`: interval -> loc + scale * sigmoid(u), greater_than(low) -> low + exp(u), real -> identity,
with the log|dx/du| terms added to the potential.  15 scalars: stays in Python.
"""
import math

import numpy as np

_SQRT2 = math.sqrt(2.0)
_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def _Phi(x):
    return 0.5 * math.erfc(-x / _SQRT2)


class Site:
    """kind: 'tn' (truncated normal), 'n' (normal), 'u' (uniform)."""

    def __init__(self, name, kind, a, b, lo=-math.inf, hi=math.inf):
        self.name, self.kind, self.a, self.b, self.lo, self.hi = name, kind, float(a), float(b), float(lo), float(hi)
        if kind == "u":
            self.lo, self.hi = float(a), float(b)
        if kind == "tn":
            za = (self.lo - self.a) / self.b if math.isfinite(self.lo) else -math.inf
            zb = (self.hi - self.a) / self.b if math.isfinite(self.hi) else math.inf
            self.log_z = math.log((1.0 if zb == math.inf else _Phi(zb)) - (0.0 if za == -math.inf else _Phi(za)))

    # ---- log density and d/dx in constrained space
    def log_prob(self, x):
        if self.kind == "u":
            return -math.log(self.hi - self.lo) if self.lo <= x <= self.hi else -math.inf
        z = (x - self.a) / self.b
        lp = -0.5 * z * z - _LOG_SQRT_2PI - math.log(self.b)
        if self.kind == "tn":
            if not (self.lo <= x <= self.hi):
                return -math.inf
            lp -= self.log_z
        return lp

    def dlog_prob(self, x):
        return 0.0 if self.kind == "u" else -(x - self.a) / (self.b * self.b)

    # ---- biject_to(support): x(u), dx/du, log|dx/du| and d log|dx/du| / du
    def forward(self, u):
        lo, hi = self.lo, self.hi
        if math.isfinite(lo) and math.isfinite(hi):
            s = 1.0 / (1.0 + math.exp(-u)) if u >= 0 else math.exp(u) / (1.0 + math.exp(u))
            x = lo + (hi - lo) * s
            dx = (hi - lo) * s * (1.0 - s)
            return x, dx, math.log(hi - lo) + _log_sigmoid(u) + _log_sigmoid(-u), 1.0 - 2.0 * s
        if math.isfinite(lo):
            e = math.exp(u)
            return lo + e, e, u, 1.0
        return u, 1.0, 0.0, 0.0

    def inverse(self, x):
        lo, hi = self.lo, self.hi
        if math.isfinite(lo) and math.isfinite(hi):
            s = (x - lo) / (hi - lo)
            return math.log(s) - math.log1p(-s)
        if math.isfinite(lo):
            return math.log(x - lo)
        return x

    def sample(self, rng):
        if self.kind == "u":
            return rng.uniform(self.lo, self.hi)
        while True:
            x = rng.normal(self.a, self.b)
            if self.lo <= x <= self.hi:
                return x


def _log_sigmoid(u):
    return -math.log1p(math.exp(-u)) if u >= 0 else u - math.log1p(math.exp(u))


# Sample sites in the order the reference declares them inside pop_cosmo_model (:368-372, :398):
# cosmo_parameters, mass_parameters, redshift_parameters, R_unit.
SITES = (
    Site("h", "tn", 0.7, 0.2, 0.35, 1.4),              # :306
    Site("Om", "tn", 0.3, 0.15, 0.0, 1.0),             # :307
    Site("w", "tn", -1.0, 0.25, -1.5, -0.5),           # :308
    Site("a", "tn", 2.35, 2.0, -1.65, 6.35),           # :282
    Site("b", "tn", 1.9, 2.0, -2.1, 5.9),              # :283
    Site("c", "tn", 4.0, 2.0, 0.0, 8.0),               # :284
    Site("mpisn", "tn", 35.0, 5.0, 20.0, 50.0),        # :286
    Site("dmbhmax", "tn", 5.0, 2.0, 0.5, 11.0),        # :287
    Site("sigma", "tn", 2.0, 2.0, 1.0),                # :289
    Site("beta", "n", 0.0, 2.0),                       # :291
    Site("log_fpl", "u", math.log(1e-3), math.log(0.5)),   # :293
    Site("lam", "tn", 2.7, 2.0, -1.3, 6.7),            # :299
    Site("dkappa", "tn", 5.6 - 2.7, 2.0, 1.0, 9.6 - 2.7),  # :300
    Site("zp", "tn", 1.9, 1.0, 0.0, 3.9),              # :302
    Site("R_unit", "n", 0.0, 1.0),                     # :398
)
SITE_NAMES = tuple(s.name for s in SITES)
NSITES = len(SITES)              # 15; the first 14 enter the likelihood
FIXED_SITES = SITES[3:]          # pop_model (fixed cosmology, intensity_models.py:313-355): no h, Om, w
LIKELIHOOD_SITES = SITE_NAMES[:14]


def theta_from_sites(x):
    """Derived kernel parameters (h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp) from the
    14 likelihood sites in SITES order (deterministics :288, :294, :301)."""
    h, Om, w, a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp = (float(v) for v in x[:14])
    return np.array([h, Om, w, a, b, c, mpisn, mpisn + dmbhmax, sigma, math.exp(log_fpl), beta, lam, lam + dkappa, zp])


def grad_sites_from_theta(g, theta):
    """Chain rule d/d(theta) -> d/d(14 likelihood sites), SITES order."""
    (gh, gOm, gw, ga, gb, gc, gmpisn, gmbhmax, gsigma, gfpl, gbeta, glam, gkappa, gzp) = (float(v) for v in g[:14])
    fpl = float(theta[9])
    return np.array([gh, gOm, gw, ga, gb, gc, gmpisn + gmbhmax, gmbhmax, gsigma, gbeta, fpl * gfpl, glam + gkappa,
                     gkappa, gzp])


def log_prior(x):
    """Sum of the 15 site log-densities (constrained space) and its gradient."""
    lp = 0.0
    g = np.zeros(NSITES)
    for i, s in enumerate(SITES):
        lp += s.log_prob(float(x[i]))
        g[i] = s.dlog_prob(float(x[i]))
    return lp, g


def constrain(u, sites=None):
    """Unconstrained vector -> (x, dx/du, sum log|dx/du|, d(sum log|dx/du|)/du).  `sites`: the model's site table
    (default: the 15 sites of pop_cosmo_model)."""
    sites = SITES if sites is None else sites
    n = len(sites)
    x = np.empty(n)
    dx = np.empty(n)
    dlj = np.empty(n)
    lj = 0.0
    for i, s in enumerate(sites):
        x[i], dx[i], l, dlj[i] = s.forward(float(u[i]))
        lj += l
    return x, dx, lj, dlj


def potential_terms(u, sites=None):
    """One pass over the sites for NUTS: (x, dx/du, log prior + log|dx/du|, d/du of that sum's explicit part).

    Equivalent to constrain() followed by log_prior(), fused and kept in plain Python floats (15 scalars: numpy
    costs more than it saves here).  Returns (x list, dx list, logp_plus_logjac, glp list, dlj list) where the
    gradient of the total with respect to u is glp*dx + dlj."""
    xs, dxs, glps, dljs = [], [], [], []
    total = 0.0
    exp, log1p, log = math.exp, math.log1p, math.log
    for s, ui in zip(SITES if sites is None else sites, u):
        ui = float(ui)
        lo, hi = s.lo, s.hi
        if lo != -math.inf and hi != math.inf:
            if ui >= 0:
                e = exp(-ui)
                sg = 1.0 / (1.0 + e)
                lsig2 = -ui - 2.0 * log1p(e)
            else:
                e = exp(ui)
                sg = e / (1.0 + e)
                lsig2 = ui - 2.0 * log1p(e)
            w = hi - lo
            x = lo + w * sg
            dx = w * sg * (1.0 - sg)
            total += log(w) + lsig2
            dlj = 1.0 - 2.0 * sg
        elif lo != -math.inf:
            e = exp(ui)
            x, dx, dlj = lo + e, e, 1.0
            total += ui
        else:
            x, dx, dlj = ui, 1.0, 0.0
        if s.kind == "u":
            total -= log(hi - lo)
            glp = 0.0
        else:
            z = (x - s.a) / s.b
            total += -0.5 * z * z - _LOG_SQRT_2PI - log(s.b) - (s.log_z if s.kind == "tn" else 0.0)
            glp = -z / s.b
        xs.append(x)
        dxs.append(dx)
        glps.append(glp)
        dljs.append(dlj)
    return xs, dxs, total, glps, dljs


def unconstrain(x):
    return np.array([s.inverse(float(x[i])) for i, s in enumerate(SITES)])


def sample_prior(rng):
    return np.array([s.sample(rng) for s in SITES])

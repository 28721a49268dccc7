"""CPU oracle for the BumpCosmology hyperlikelihood hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A float64 restatement (torch, CPU) of the reference's `pop_cosmo_model` likelihood body and everything
it evaluates, with reverse-mode gradients from `torch.autograd`:

    /root/reference/src/scripts/intensity_models.py:13-54    helpers (core->remnant map, turn-on)
    /root/reference/src/scripts/intensity_models.py:56-111   LogDNDMPISN   (tabulated pile-up mass function)
    /root/reference/src/scripts/intensity_models.py:113-151  LogDNDM       (+ power-law tail, normalisation)
    /root/reference/src/scripts/intensity_models.py:153-173  LogDNDV       (Madau-Dickinson-like rate)
    /root/reference/src/scripts/intensity_models.py:175-210  LogDNDMDQDV   (joint density, pairing beta)
    /root/reference/src/scripts/intensity_models.py:212-273  FlatwCDMCosmology (1024-pt tables, interp accessors)
    /root/reference/src/scripts/intensity_models.py:357-401  pop_cosmo_model likelihood body
    /root/reference/src/scripts/utils.py:3-8                 jnp_cumtrapz

Third-party arithmetic that is NOT in the reference tree (jax / jaxlib, unpinned in environment.yml) is
restated from its documented semantics: `jnp.interp` (searchsorted side='right' clipped to [1, n-1],
end-clamped, dx~0 guard), `jnp.logaddexp`, `jax.scipy.special.logsumexp`, `jnp.linspace`.

PARITY STATUS: the reference has no tests, golden vectors or recorded outputs, and JAX cannot
be installed here, so parity is not pinned by upstream fixtures.  It IS pinned to the reference *source*:
`oracle/run_reference.py` executes the unmodified `intensity_models.py` on a torch-backed stand-in for
jax/numpyro, and `tests/golden/*.npz` (minted by `tests/golden/make_golden.py` from that run) are what this
restatement is checked against in `tests/test_oracle.py`.  What remains unpinned is the third-party layer
itself (real jax.numpy / XLA:CPU rounding).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference` legs may import
this module.  The product (`bumpcosmology_b200/`) never does and has no CPU fallback.

theta order (the *derived* parameters the density objects receive, intensity_models.py:370-376):
    (h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp)   [+ optional wa]
"""
import math

import numpy as np
import torch

F = torch.float64
THETA_NAMES = ("h", "Om", "w", "a", "b", "c", "mpisn", "mbhmax", "sigma", "fpl", "beta", "lam", "kappa", "zp")
NTHETA = len(THETA_NAMES)

# constants (file:line in intensity_models.py)
MBH_MIN = 5.0          # :13
MTR = 20.0             # :41
TURNON_WIDTH = 0.05    # :45
N_M = 256              # :92
MIN_BH_MASS = 3.0      # :97
MIN_CO_MASS = 1.0      # :98
MREF, QREF, ZREF = 30.0, 1.0, 0.0   # :129,191-193
ZMAX, NINTERP = 100.0, 1024         # :220-221
C_H100_GPC = 2.99792                # :239
NEG_INF = float("-inf")


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))


def linspace(start, stop, num):
    """jnp.linspace semantics: start*(1-s)+stop*s, s=i/(num-1), endpoint pinned; differentiable."""
    start, stop = _t(start), _t(stop)
    s = torch.arange(num, dtype=F) / (num - 1)
    out = start * (1 - s) + stop * s
    return torch.cat([out[:-1], stop.reshape(1)])


def interp(x, xp, fp):
    """jnp.interp semantics; gradients flow to x, xp and fp, not to the index."""
    x, xp, fp = _t(x), _t(xp), _t(fp)
    n = xp.shape[0]
    i = torch.clamp(torch.searchsorted(xp.detach(), x.detach().contiguous(), right=True), 1, n - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    eps = float(np.spacing(np.finfo(np.float64).eps))
    dx0 = dx.abs() <= eps
    f = torch.where(dx0, fp[i - 1], fp[i - 1] + (delta / torch.where(dx0, torch.ones_like(dx), dx)) * df)
    f = torch.where(x < xp[0], fp[0], f)
    f = torch.where(x > xp[-1], fp[-1], f)
    return f


def cumtrapz(ys, xs):
    """utils.py:3-8."""
    return torch.cat([torch.zeros(1, dtype=F), torch.cumsum(0.5 * torch.diff(xs) * (ys[:-1] + ys[1:]), 0)])


# --------------------------------------------------------------------------- cosmology (:212-273)
class Cosmology:
    """FlatwCDMCosmology.__post_init__ (:229-235); `wa` is the CPL extension of BASELINE.json's last
    config (no reference counterpart; identical to the reference at wa = 0)."""

    def __init__(self, h, Om, w, wa=None):
        self.h, self.Om, self.w, self.wa = _t(h), _t(Om), _t(w), (None if wa is None else _t(wa))
        self.zinterp = torch.expm1(torch.as_tensor(np.linspace(np.log(1), np.log(1 + ZMAX), NINTERP)))
        self.dH = C_H100_GPC / self.h
        E = self.E(self.zinterp)
        self.dcinterp = self.dH * cumtrapz(1 / E, self.zinterp)
        self.dlinterp = self.dcinterp * (1 + self.zinterp)
        self.ddlinterp = self.dcinterp + self.dH * (1 + self.zinterp) / E
        self.dvcinterp = 4 * np.pi * self.dcinterp * self.dcinterp * self.dH / E

    def E(self, z):  # :253-256
        opz = 1 + z
        opz3 = opz * opz * opz
        if self.wa is None:
            de = opz ** (3 * (1 + self.w))
        else:
            de = opz ** (3 * (1 + self.w + self.wa)) * torch.exp(-3 * self.wa * z / opz)
        return torch.sqrt(self.Om * opz3 + (1 - self.Om) * de)

    def z_of_dL(self, dl):  # :272-273
        return interp(dl, self.dlinterp, self.zinterp)

    def dVCdz(self, z):  # :264-265
        return interp(z, self.zinterp, self.dvcinterp)

    def ddL_dz(self, z):  # :267-268
        return interp(z, self.zinterp, self.ddlinterp)


# --------------------------------------------------------------------------- mass function (:15-151)
def mean_mbh_from_mco(mco, mpisn, mbhmax):  # :15-25
    a = 1 / (4 * (mpisn - mbhmax))
    mcomax = 2 * mbhmax - mpisn
    return torch.where(mco < mpisn, mco, mbhmax + a * (mco - mcomax) ** 2)


def largest_mco(mpisn, mbhmax):  # :27-30
    mcomax = 2 * mbhmax - mpisn
    return mcomax + torch.sqrt(4 * mbhmax * (mbhmax - mpisn))


def log_dNdmCO(mco, a, b):  # :32-43
    x = mco / MTR
    return torch.where(mco < MTR, -a * torch.log(x), -b * torch.log(x))


def log_smooth_turnon(m, mmin, width=TURNON_WIDTH):  # :45-54
    dm = mmin * width
    return math.log(2) - torch.log1p(torch.exp(-(m - mmin) / dm))


class PISNTable:
    """LogDNDMPISN (:96-111)."""

    def __init__(self, a, b, mpisn, mbhmax, sigma):
        a, b, mpisn, mbhmax, sigma = map(_t, (a, b, mpisn, mbhmax, sigma))
        max_bh_mass = mbhmax + 7 * sigma
        max_co_mass = largest_mco(mpisn, mbhmax)
        mbh = linspace(MIN_BH_MASS, max_bh_mass, N_M)
        mco = linspace(MIN_CO_MASS, max_co_mass, N_M)
        log_wts = (log_dNdmCO(mco[None, :], a, b)
                   - 0.5 * ((mbh[:, None] - mean_mbh_from_mco(mco[None, :], mpisn, mbhmax)) / sigma) ** 2
                   - np.log(np.sqrt(2 * np.pi)) - torch.log(sigma))
        log_trapz = (np.log(0.5) + torch.logaddexp(log_wts[:, 1:], log_wts[:, :-1])
                     + torch.log(torch.diff(mco[None, :], dim=1)))
        self.log_dN_grid = torch.logsumexp(log_trapz, dim=1)
        self.mbh_grid = mbh

    def __call__(self, m):
        return interp(m, self.mbh_grid, self.log_dN_grid)


class MassFunction:
    """LogDNDM (:134-151)."""

    def __init__(self, a, b, c, mpisn, mbhmax, sigma, fpl):
        self.c, self.mbhmax = _t(c), _t(mbhmax)
        self.pisn = PISNTable(a, b, mpisn, mbhmax, sigma)
        self.log_pl_norm = torch.log(_t(fpl)) + self.pisn(self.mbhmax)
        self.log_norm = torch.zeros((), dtype=F)
        self.log_norm = -(self(torch.tensor(MREF, dtype=F)) + math.log(MREF))

    def __call__(self, m):
        m = _t(m)
        ld = self.pisn(m)
        ninf = torch.full_like(ld, NEG_INF)
        ld = torch.where(m <= self.pisn.mbh_grid[0], ninf, ld)
        ld = torch.where(m >= self.pisn.mbh_grid[-1], ninf, ld)
        ld = torch.logaddexp(ld, -self.c * torch.log(m / self.mbhmax) + self.log_pl_norm
                             + log_smooth_turnon(m, self.mbhmax))
        ld = torch.where(m < MBH_MIN, ninf, ld)
        return ld + self.log_norm


class Rate:
    """LogDNDV (:167-173)."""

    def __init__(self, lam, kappa, zp):
        self.lam, self.kappa, self.zp = _t(lam), _t(kappa), _t(zp)
        self.log_norm = torch.zeros((), dtype=F)
        self.log_norm = -self(torch.tensor(ZREF, dtype=F))

    def __call__(self, z):
        z = _t(z)
        return self.lam * torch.log1p(z) - torch.log1p(((1 + z) / (1 + self.zp)) ** self.kappa) + self.log_norm


class JointDensity:
    """LogDNDMDQDV (:198-210)."""

    def __init__(self, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp):
        self.beta = _t(beta)
        self.log_dndm = MassFunction(a, b, c, mpisn, mbhmax, sigma, fpl)
        self.log_dndv = Rate(lam, kappa, zp)

    def __call__(self, m1, q, z):
        m1, q, z = _t(m1), _t(q), _t(z)
        m2 = q * m1
        mt = m1 + m2
        return (self.log_dndm(m1) + self.log_dndm(m2) + self.beta * torch.log(mt / (MREF * (1 + QREF)))
                + torch.log(m1) + self.log_dndv(z))


# --------------------------------------------------------------------------- likelihood body (:357-401)
def split_theta(theta):
    return {k: theta[i] for i, k in enumerate(THETA_NAMES)}


def build_model(theta, wa=None):
    p = split_theta(theta)
    cosmo = Cosmology(p["h"], p["Om"], p["w"], wa)
    log_dN = JointDensity(p["a"], p["b"], p["c"], p["mpisn"], p["mbhmax"], p["sigma"], p["fpl"], p["beta"],
                          p["lam"], p["kappa"], p["zp"])
    return cosmo, log_dN


def log_weights(cosmo, log_dN, m1_det, q, dl, pdraw):
    """:378-381 / :385-388."""
    m1_det, q, dl, pdraw = map(_t, (m1_det, q, dl, pdraw))
    zs = cosmo.z_of_dL(dl)
    m1s = m1_det / (1 + zs)
    return (log_dN(m1s, q, zs) - 2 * torch.log1p(zs) + torch.log(cosmo.dVCdz(zs))
            - torch.log(cosmo.ddL_dz(zs)) - torch.log(pdraw))


def event_terms(cosmo, log_dN, m1s_det, qs, dls, pdraw):
    """Per-event log-mean-exp (:382) and Neff (:401). Returns (log_like[nobs], neff[nobs])."""
    lw = log_weights(cosmo, log_dN, m1s_det, qs, dls, pdraw)
    nsamp = lw.shape[1]
    lse = torch.logsumexp(lw, dim=1)
    neff = torch.exp(2 * lse - torch.logsumexp(2 * lw, dim=1))
    return lse - math.log(nsamp), neff


def selection_terms(cosmo, log_dN, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw):
    """:385-394. Returns (log_mu_sel, log_mu2, neff_sel)."""
    lw = log_weights(cosmo, log_dN, m1s_det_sel, qs_sel, dls_sel, pdraw_sel)
    log_nd = math.log(float(Ndraw))
    log_mu = torch.logsumexp(lw, 0) - log_nd
    log_mu2 = torch.logsumexp(2 * lw, 0) - 2 * log_nd
    log_s2 = log_mu2 + torch.log1p(-torch.exp(2 * log_mu - log_nd - log_mu2))
    return log_mu, log_mu2, torch.exp(2 * log_mu - log_s2)


def evaluate(theta, data, grad=True, wa=None, event_chunk=None):
    """One evaluation of the hot path = what the C-ABI `bump_eval` replaces.

    theta: 14 floats (kernel order).  data: the 9 positional arguments of pop_cosmo_model.
    Returns a dict: loglike (sum over events, the 'loglike' factor :383), log_mu_sel (:389; the
    'selfactor' is -nobs*log_mu_sel :390), log_mu2, neff_sel, neff[nobs], and (grad=True)
    dloglike[14], dlog_mu_sel[14] (+ `dloglike_dwa`, `dlog_mu_sel_dwa` when wa is given).
    event_chunk bounds temporaries for big catalogs (gradients are accumulated chunk by chunk).
    """
    m1s_det, qs, dls, pdraw, m1s_sel, qs_sel, dls_sel, pdraw_sel, Ndraw = data
    m1s_det, qs, dls, pdraw = (np.asarray(x, dtype=np.float64) for x in (m1s_det, qs, dls, pdraw))
    nobs = m1s_det.shape[0]
    th = torch.tensor(np.asarray(theta, dtype=np.float64)[:NTHETA], requires_grad=grad)
    wa_t = None if wa is None else torch.tensor(float(wa), dtype=F, requires_grad=grad)
    leaves = [th] + ([] if wa_t is None else [wa_t])

    def model():
        return build_model(th, wa_t)

    ng = len(leaves)
    g_ll = [torch.zeros_like(v) for v in leaves]
    loglike = 0.0
    neff = np.empty(nobs)
    step = nobs if not event_chunk else int(event_chunk)
    for lo in range(0, nobs, max(step, 1)):
        hi = min(nobs, lo + step)
        cosmo, log_dN = model()
        ll, ne = event_terms(cosmo, log_dN, m1s_det[lo:hi], qs[lo:hi], dls[lo:hi], pdraw[lo:hi])
        s = ll.sum()
        if grad:
            gs = torch.autograd.grad(s, leaves)
            for k in range(ng):
                g_ll[k] += gs[k]
        loglike += float(s.detach())
        neff[lo:hi] = ne.detach().numpy()
    cosmo, log_dN = model()
    log_mu, log_mu2, neff_sel = selection_terms(cosmo, log_dN, m1s_sel, qs_sel, dls_sel, pdraw_sel, Ndraw)
    out = {"loglike": loglike, "log_mu_sel": float(log_mu.detach()), "log_mu2": float(log_mu2.detach()),
           "neff_sel": float(neff_sel.detach()), "neff": neff, "nobs": nobs,
           "selfactor": -nobs * float(log_mu.detach()), "logl": loglike - nobs * float(log_mu.detach())}
    if grad:
        g_mu = torch.autograd.grad(log_mu, leaves)
        out["dloglike"] = g_ll[0].numpy().copy()
        out["dlog_mu_sel"] = g_mu[0].numpy().copy()
        out["dlogl"] = out["dloglike"] - nobs * out["dlog_mu_sel"]
        if wa_t is not None:
            out["dloglike_dwa"] = float(g_ll[1])
            out["dlog_mu_sel_dwa"] = float(g_mu[1])
    return out


def evaluate_fixed(theta, data, dvdzdt_interp, grad=True):
    """Fixed-cosmology variant = the reference's `pop_model` (intensity_models.py:313-355).

    theta: the same 14-vector (entries 0..2 = h, Om, w are ignored); data: (m1s, qs, zs, pdraw, m1s_sel, qs_sel,
    zs_sel, pdraw_sel, Ndraw) in the SOURCE frame; dvdzdt_interp: the theta-independent table of :325 on
    zinterp = expm1(linspace(log1p(0), log1p(100), 1024)) (:324)."""
    m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw = data
    th = torch.tensor(np.asarray(theta, dtype=np.float64)[:NTHETA], requires_grad=grad)
    p = split_theta(th)
    log_dN = JointDensity(p["a"], p["b"], p["c"], p["mpisn"], p["mbhmax"], p["sigma"], p["fpl"], p["beta"],
                          p["lam"], p["kappa"], p["zp"])
    zinterp = torch.expm1(torch.as_tensor(np.linspace(np.log1p(0), np.log1p(ZMAX), NINTERP)))
    tab = _t(dvdzdt_interp)

    def lw(m1, q, z, pd):   # :332 / :336
        m1, q, z, pd = map(_t, (m1, q, z, pd))
        return log_dN(m1, q, z) + torch.log(interp(z, zinterp, tab)) - torch.log(pd)

    w = lw(m1s, qs, zs, pdraw)
    nobs, nsamp = w.shape
    lse = torch.logsumexp(w, dim=1)
    loglike = (lse - math.log(nsamp)).sum()
    neff = torch.exp(2 * lse - torch.logsumexp(2 * w, dim=1))
    ws = lw(m1s_sel, qs_sel, zs_sel, pdraw_sel)
    log_nd = math.log(float(Ndraw))
    log_mu = torch.logsumexp(ws, 0) - log_nd
    log_mu2 = torch.logsumexp(2 * ws, 0) - 2 * log_nd
    log_s2 = log_mu2 + torch.log1p(-torch.exp(2 * log_mu - log_nd - log_mu2))
    out = {"loglike": float(loglike.detach()), "log_mu_sel": float(log_mu.detach()), "log_mu2": float(log_mu2.detach()),
           "neff_sel": float(torch.exp(2 * log_mu - log_s2).detach()), "neff": neff.detach().numpy().copy(),
           "nobs": nobs, "selfactor": -nobs * float(log_mu.detach())}
    if grad:
        out["dloglike"] = torch.autograd.grad(loglike, th, retain_graph=True)[0].numpy().copy()
        out["dlog_mu_sel"] = torch.autograd.grad(log_mu, th)[0].numpy().copy()
    return out


def tables(theta, wa=None):
    """theta-dependent tables (for unit-level parity of the prologue kernels)."""
    with torch.no_grad():
        cosmo, log_dN = build_model(torch.as_tensor(np.asarray(theta, dtype=np.float64)), None if wa is None
                                    else torch.tensor(float(wa), dtype=F))
    n = lambda t: t.numpy().copy()  # noqa: E731
    mf = log_dN.log_dndm
    return {"zinterp": n(cosmo.zinterp), "dcinterp": n(cosmo.dcinterp), "dlinterp": n(cosmo.dlinterp),
            "ddlinterp": n(cosmo.ddlinterp), "dvcinterp": n(cosmo.dvcinterp),
            "mbh_grid": n(mf.pisn.mbh_grid), "log_dN_grid": n(mf.pisn.log_dN_grid),
            "log_pl_norm": float(mf.log_pl_norm), "log_norm": float(mf.log_norm),
            "rate_log_norm": float(log_dN.log_dndv.log_norm)}


def table_jacobians(theta):
    """d(table)/d(theta) by autograd, for checking the prologue's forward-mode tangent tables."""
    th = torch.tensor(np.asarray(theta, dtype=np.float64)[:NTHETA], requires_grad=True)
    cosmo, log_dN = build_model(th)
    mf = log_dN.log_dndm
    outs = {"dlinterp": cosmo.dlinterp, "ddlinterp": cosmo.ddlinterp, "dvcinterp": cosmo.dvcinterp,
            "log_dN_grid": mf.pisn.log_dN_grid, "log_pl_norm": mf.log_pl_norm.reshape(1),
            "log_norm": mf.log_norm.reshape(1), "rate_log_norm": log_dN.log_dndv.log_norm.reshape(1)}
    jac = {}
    for k, v in outs.items():
        rows = []
        for i in range(v.shape[0]):
            (g,) = torch.autograd.grad(v[i], th, retain_graph=True, allow_unused=True)
            rows.append(np.zeros(NTHETA) if g is None else g.numpy().copy())
        jac[k] = np.array(rows)
    return jac


# --------------------------------------------------------------------------- sample-site chain rule
SAMPLE_SITES = ("h", "Om", "w", "a", "b", "c", "mpisn", "dmbhmax", "sigma", "beta", "log_fpl",
                "lam", "dkappa", "zp")


def theta_from_sites(s):
    """Derived parameters (:288,294,301)."""
    return np.array([s["h"], s["Om"], s["w"], s["a"], s["b"], s["c"], s["mpisn"], s["mpisn"] + s["dmbhmax"],
                     s["sigma"], math.exp(s["log_fpl"]), s["beta"], s["lam"], s["lam"] + s["dkappa"], s["zp"]])


def grad_sites_from_theta(g, theta):
    """Chain rule from d/d(theta) to d/d(sample sites) in SAMPLE_SITES order."""
    g = dict(zip(THETA_NAMES, g))
    fpl = theta[THETA_NAMES.index("fpl")]
    return np.array([g["h"], g["Om"], g["w"], g["a"], g["b"], g["c"], g["mpisn"] + g["mbhmax"], g["mbhmax"],
                     g["sigma"], g["beta"], fpl * g["fpl"], g["lam"] + g["kappa"], g["kappa"], g["zp"]])

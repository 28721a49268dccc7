"""Seeded synthetic catalogs in the shapes BASELINE.json names (GWTC-3, O4, O5).

There is no network and the reference ships no data (its inputs are Zenodo downloads), so the events'
detector-frame posterior samples and the found-injection set are synthesised following the reference's
own mock recipes:

* true population: `mock_injections.py:83-104,137-157` (Madau-Dickinson-like z pdf times dVc/dz/(1+z),
  m1 ~ m^-2.35 on [5, 500], total mass ~ M^-2 on [m1+5, 2 m1]);
* observation noise and posterior samples: `mock_observations.py:41-48` and
  `weighting.draw_mock_samples` (`weighting.py:182-215`, detector-frame branch): Gaussian in
  log Mc_det, q (truncated to (0, 1]) and log d_L with widths 0.05/0.07/0.2 * 20/rho,
  `pdraw = 1/(m1_det d_L)`;
* detector-frame conversion and Jacobian of the injections: `run_cosmo_fit.py:27-30`,
  `weighting.py:173-180`, with a fixed flat LCDM (h=0.6766, Om=0.30966) standing in for astropy's Planck18.

Layout matches what `run_cosmo_fit.py:32-49` hands to `pop_cosmo_model`: four `[nobs, nsamp]` float64
arrays, four `[nsel]` float64 arrays and the scalar `Ndraw`.
"""
from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (nobs, nsamp, nsel, seed, zmax of the true events)
    "tiny": (5, 64, 512, 20230, 1.5),
    "small": (12, 256, 4096, 20229, 1.5),
    "gwtc3": (69, 4096, 200_000, 20231, 1.5),
    "o4": (300, 8192, 1_000_000, 20232, 1.5),
    "o5": (5000, 10000, 10_000_000, 20233, 3.0),
    # GWTC-3 shape for sampler runs: every sample and injection keeps its source-frame m2 well above the model's
    # hard cut at mbh_min = 5 (see make_catalog: mass_floor)
    "gwtc3_nuts": (69, 4096, 200_000, 20234, 1.5),
}
MASS_FLOOR = {"gwtc3_nuts": 8.0}

_H_FID, _OM_FID = 0.6766, 0.30966
_C_H100_GPC = 2.99792  # same constant as intensity_models.py:239


@dataclass
class Catalog:
    m1s_det: np.ndarray
    qs: np.ndarray
    dls: np.ndarray
    pdraw: np.ndarray
    m1s_det_sel: np.ndarray
    qs_sel: np.ndarray
    dls_sel: np.ndarray
    pdraw_sel: np.ndarray
    Ndraw: float
    name: str = "custom"

    def as_args(self):
        """The 9 positional arguments of `pop_cosmo_model` (intensity_models.py:357)."""
        return (self.m1s_det, self.qs, self.dls, self.pdraw, self.m1s_det_sel, self.qs_sel, self.dls_sel,
                self.pdraw_sel, self.Ndraw)

    @property
    def nobs(self):
        return self.m1s_det.shape[0]

    @property
    def nsamp(self):
        return self.m1s_det.shape[1]

    @property
    def nsel(self):
        return self.m1s_det_sel.shape[0]

    @property
    def n_elements(self):
        return self.nobs * self.nsamp + self.nsel


class _FiducialCosmology:
    """Flat LCDM distance tables on a 4096-point grid uniform in log(1+z)."""

    def __init__(self, zmax=20.0, n=4096):
        self.z = np.expm1(np.linspace(0.0, np.log1p(zmax), n))
        self.dH = _C_H100_GPC / _H_FID
        self.E = np.sqrt(_OM_FID * (1 + self.z) ** 3 + (1 - _OM_FID))
        inv = 1.0 / self.E
        self.dc = self.dH * np.concatenate(([0.0], np.cumsum(0.5 * np.diff(self.z) * (inv[:-1] + inv[1:]))))
        self.dl = self.dc * (1 + self.z)
        self.ddl = self.dc + self.dH * (1 + self.z) / self.E
        self.dvc = 4 * np.pi * self.dc ** 2 * self.dH / self.E

    def f(self, z, tab):
        return np.interp(z, self.z, tab)


class _PowerLaw:
    """p(x) ~ x^-alpha on [a, b] (mock_injections.py:117-133)."""

    def __init__(self, alpha, a, b):
        self.alpha, self.a, self.b = alpha, a, b
        self.lognorm = np.log((a ** (1 - alpha) - b ** (1 - alpha)) / (alpha - 1))

    def pdf(self, x):
        return np.exp(-self.alpha * np.log(x) - self.lognorm)

    def icdf(self, c):
        al = self.alpha
        return (self.a ** (1 - al) * (1 - c) + self.b ** (1 - al) * c) ** (1 / (1 - al))


class _ZPDF:
    def __init__(self, cosmo, zmax):
        self.cosmo = cosmo
        self.zg = np.expm1(np.linspace(0.0, np.log1p(zmax), 2048))
        un = self._unnorm(self.zg)
        cdf = np.concatenate(([0.0], np.cumsum(0.5 * np.diff(self.zg) * (un[:-1] + un[1:]))))
        self.norm = 1.0 / cdf[-1]
        self.cdf = cdf * self.norm

    def _unnorm(self, z):
        return (1 + z) ** 2.7 / (1 + ((1 + z) / 2.9) ** 5.6) * self.cosmo.f(z, self.cosmo.dvc) / (1 + z)

    def pdf(self, z):
        return self.norm * self._unnorm(z)

    def icdf(self, c):
        return np.interp(c, self.cdf, self.zg)


def _draw_population(rng, n, zpdf, m1_lo=5.0, m2_lo=5.0):
    z = zpdf.icdf(rng.uniform(size=n))
    mpdf = _PowerLaw(2.35, m1_lo, 500.0)
    m1 = mpdf.icdf(rng.uniform(size=n))
    mt_pdf = _PowerLaw(2.0, m1 + m2_lo, 2 * m1)
    mt = mt_pdf.icdf(rng.uniform(size=n))
    q = np.minimum((mt - m1) / m1, 1.0)
    pdraw_mqz = mpdf.pdf(m1) * (mt_pdf.pdf(mt) * m1) * zpdf.pdf(z)
    return m1, q, z, pdraw_mqz


def make_catalog(name="gwtc3", nobs=None, nsamp=None, nsel=None, seed=None, zmax=None, chunk_events=256,
                 mass_floor=None):
    """Build a seeded synthetic catalog. `name` picks a shape from SHAPES; explicit sizes override it.

    mass_floor (default: MASS_FLOOR.get(name)): if set, posterior samples are redrawn until their source-frame m2
    (at the fiducial cosmology) is >= mass_floor, and injections are drawn with m1, m2 >= mass_floor.  The
    reference's mass function is cut hard at mbh_min = 5 (intensity_models.py:149): samples that cross the cut as
    (h, Om, w) move make logL discontinuous while its gradient (JAX's as well as ours) ignores the jumps, which
    stalls any HMC sampler.  Real PE samples sit away from the cut; the sampler workload mimics that."""
    d_nobs, d_nsamp, d_nsel, d_seed, d_zmax = SHAPES.get(name, SHAPES["gwtc3"])
    nobs = d_nobs if nobs is None else nobs
    nsamp = d_nsamp if nsamp is None else nsamp
    nsel = d_nsel if nsel is None else nsel
    seed = d_seed if seed is None else seed
    zmax = d_zmax if zmax is None else zmax
    mass_floor = MASS_FLOOR.get(name) if mass_floor is None else mass_floor
    rng = np.random.default_rng(seed)
    cosmo = _FiducialCosmology()

    # ---- true events: heavier lower mass bound so that every event keeps finite-weight samples
    # (the reference rejects events whose median m2 < mbh_min = 5, weighting.py:88-89).
    zpdf_evt = _ZPDF(cosmo, zmax)
    m1_t = np.empty(nobs)
    q_t = np.empty(nobs)
    z_t = np.empty(nobs)
    filled = 0
    while filled < nobs:
        m1, q, z, _ = _draw_population(rng, 4 * (nobs - filled) + 16, zpdf_evt, m1_lo=8.0)
        ok = (q * m1 >= (7.0 if mass_floor is None else mass_floor + 6.0)) & (m1 <= 120.0)
        k = min(int(ok.sum()), nobs - filled)
        m1_t[filled:filled + k] = m1[ok][:k]
        q_t[filled:filled + k] = q[ok][:k]
        z_t[filled:filled + k] = z[ok][:k]
        filled += k
    rho = rng.uniform(10.0, 30.0, size=nobs)
    s_mc, s_q, s_dl = 0.05 * 20 / rho, 0.07 * 20 / rho, 0.2 * 20 / rho
    mc_det_t = m1_t * (1 + z_t) * q_t ** 0.6 / (1 + q_t) ** 0.2
    dl_t = cosmo.f(z_t, cosmo.dl)
    log_mc_obs = np.log(mc_det_t) + s_mc * rng.standard_normal(nobs)
    q_obs = np.clip(q_t + s_q * rng.standard_normal(nobs), 0.2, 1.0)
    log_dl_obs = np.log(dl_t) + s_dl * rng.standard_normal(nobs)

    m1s_det = np.empty((nobs, nsamp))
    qs = np.empty((nobs, nsamp))
    dls = np.empty((nobs, nsamp))
    for lo in range(0, nobs, chunk_events):
        hi = min(nobs, lo + chunk_events)
        shp = (hi - lo, nsamp)
        log_mcs = log_mc_obs[lo:hi, None] + s_mc[lo:hi, None] * rng.standard_normal(shp)
        qq = q_obs[lo:hi, None] + s_q[lo:hi, None] * rng.standard_normal(shp)
        bad = (qq <= 0) | (qq > 1)
        while bad.any():  # truncated normal by redraw, weighting.py:188-191
            redraw = (np.broadcast_to(q_obs[lo:hi, None], shp)[bad]
                      + np.broadcast_to(s_q[lo:hi, None], shp)[bad] * rng.standard_normal(int(bad.sum())))
            qq[bad] = redraw
            bad = (qq <= 0) | (qq > 1)
        log_dls = log_dl_obs[lo:hi, None] + s_dl[lo:hi, None] * rng.standard_normal(shp)
        if mass_floor is not None:   # redraw samples whose source-frame m2 (fiducial cosmology) is below the floor
            for _ in range(200):
                m1d = np.exp(log_mcs) / (qq ** 0.6 / (1 + qq) ** 0.2)
                zz = np.interp(np.exp(log_dls), cosmo.dl, cosmo.z)
                bad = qq * m1d / (1 + zz) < mass_floor
                if not bad.any():
                    break
                nb = int(bad.sum())
                b_mc = np.broadcast_to(log_mc_obs[lo:hi, None], shp)[bad]
                b_smc = np.broadcast_to(s_mc[lo:hi, None], shp)[bad]
                b_q = np.broadcast_to(q_obs[lo:hi, None], shp)[bad]
                b_sq = np.broadcast_to(s_q[lo:hi, None], shp)[bad]
                b_dl = np.broadcast_to(log_dl_obs[lo:hi, None], shp)[bad]
                b_sdl = np.broadcast_to(s_dl[lo:hi, None], shp)[bad]
                log_mcs[bad] = b_mc + b_smc * rng.standard_normal(nb)
                qn = b_q + b_sq * rng.standard_normal(nb)
                qq[bad] = np.where((qn > 0) & (qn <= 1), qn, qq[bad])
                log_dls[bad] = b_dl + b_sdl * rng.standard_normal(nb)
        qs[lo:hi] = qq
        m1s_det[lo:hi] = np.exp(log_mcs) / (qq ** 0.6 / (1 + qq) ** 0.2)
        dls[lo:hi] = np.exp(log_dls)
    pdraw = 1.0 / (m1s_det * dls)  # weighting.py:214

    # ---- found injections
    zpdf_inj = _ZPDF(cosmo, 3.5)
    fl = 5.0 if mass_floor is None else mass_floor
    m1, q, z, pdraw_mqz = _draw_population(rng, nsel, zpdf_inj, m1_lo=fl, m2_lo=fl)
    jac = 1.0 / (1 + z) / (cosmo.f(z, cosmo.dc) + (1 + z) * cosmo.dH / cosmo.f(z, cosmo.E))  # weighting.py:180
    return Catalog(
        m1s_det=m1s_det, qs=qs, dls=dls, pdraw=pdraw,
        m1s_det_sel=m1 * (1 + z), qs_sel=q, dls_sel=cosmo.f(z, cosmo.dl), pdraw_sel=pdraw_mqz * jac,
        Ndraw=float(10 * nsel), name=name,
    )


# Default hyper-parameters: `weighting.py:11-24` (the reference's "reasonable fit to O3a") plus a
# Planck-like flat LCDM point.  Order is the kernel's theta order.
THETA_NAMES = ("h", "Om", "w", "a", "b", "c", "mpisn", "mbhmax", "sigma", "fpl", "beta", "lam", "kappa", "zp")
THETA_DEFAULT = np.array([0.7, 0.3, -1.0, 1.8, -0.71, 2.9, 31.0, 36.0, 2.3, 0.21, -2.2, 4.7, 7.0, 3.0])


def draw_prior_thetas(n, seed=7):
    """Seeded draws of theta from the reference's priors (intensity_models.py:281-311), by rejection
    from the untruncated normals; returned in kernel order (derived mbhmax, fpl, kappa)."""
    rng = np.random.default_rng(seed)

    def tn(mu, sd, lo, hi):
        while True:
            x = rng.normal(mu, sd)
            if lo <= x <= hi:
                return x

    out = np.empty((n, 14))
    for i in range(n):
        a = tn(2.35, 2, -1.65, 6.35)
        b = tn(1.9, 2, -2.1, 5.9)
        c = tn(4, 2, 0, 8)
        mpisn = tn(35.0, 5.0, 20.0, 50.0)
        dmbhmax = tn(5.0, 2.0, 0.5, 11.0)
        sigma = tn(2, 2, 1, np.inf)
        beta = rng.normal(0, 2)
        log_fpl = rng.uniform(np.log(1e-3), np.log(0.5))
        lam = tn(2.7, 2.0, -1.3, 6.7)
        dkappa = tn(2.9, 2.0, 1, 6.9)
        zp = tn(1.9, 1, 0, 3.9)
        h = tn(0.7, 0.2, 0.35, 1.4)
        Om = tn(0.3, 0.15, 0, 1)
        w = tn(-1, 0.25, -1.5, -0.5)
        out[i] = (h, Om, w, a, b, c, mpisn, mpisn + dmbhmax, sigma, np.exp(log_fpl), beta, lam, lam + dkappa, zp)
    return out

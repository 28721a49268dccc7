"""Host-side mirror of the reference's model interface for the hot path.

`pop_cosmo_model(m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw)` keeps the positional
signature of /root/reference/src/scripts/intensity_models.py:357 and returns a `PopCosmoModel`: the catalog is
uploaded to HBM once, and every call with sample-site values yields the numpyro-visible names of the reference
model — factors `loglike`, `selfactor` (:383, :390), deterministics `mbhmax`, `fpl`, `kappa`, `neff_sel`, `R`,
`neff`, `mdNdmdVdt_fixed_qz`, `dNdqdVdt_fixed_mz`, `dNdVdt_fixed_mq`, `hz` (:288-301, :394-406) — plus what JAX autodiff
gives numpyro implicitly: the gradient of the factors with respect to the 14 sample sites.

The likelihood and its gradient come from the CUDA library (no CPU fallback).  Only the 15-scalar prior /
transform arithmetic (priors.py) and the output-only 128-point diagnostic curves are computed on the host.
"""
import math

import numpy as np

from . import priors
from .likelihood import Hyperlikelihood, ShardedHyperlikelihood

# intensity_models.py:275-279
coords = {
    "m_grid": np.exp(np.linspace(np.log(5), np.log(150), 128)),
    "q_grid": np.linspace(0, 1, 129)[1:],
    "z_grid": np.expm1(np.linspace(np.log1p(0), np.log1p(3), 128)),
}
MREF, QREF, ZREF = 30.0, 1.0, 0.0       # :129, :191-193
SAMPLE_SITES = priors.SITE_NAMES           # 15 sites incl. R_unit
LIKELIHOOD_SITES = priors.LIKELIHOOD_SITES


class PopCosmoModel:
    """The reference model bound to its data (what numpyro holds after `mcmc.run(key, *data)`)."""

    def __init__(self, m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, device=0,
                 distributed=False, exchange="p2p"):
        data = (m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw)
        if distributed:
            self.like = ShardedHyperlikelihood(data, device=device, exchange=exchange)
            self._local = self.like.local
        else:
            self.like = self._local = Hyperlikelihood(*data, device=device)
        self.n_evals = 0

    def clone(self):
        """The same model on the same resident catalog with its own evaluation state: one per NUTS chain (the
        reference's 4 chains share one data set, run_cosmo_fit.py:46-49).  Single-rank models only."""
        if self.like is not self._local:
            raise TypeError("clone() is for single-rank models (a sharded model's ranks step in lock step)")
        new = object.__new__(PopCosmoModel)
        new.like = new._local = self.like.clone()
        new.n_evals = 0
        return new

    # ---- site handling
    @staticmethod
    def _site_vector(sites):
        if isinstance(sites, dict):
            return np.array([float(sites.get(k, 0.0)) for k in SAMPLE_SITES])
        x = np.asarray(sites, dtype=np.float64).ravel()
        if x.shape[0] == 14:
            x = np.concatenate([x, [0.0]])
        if x.shape[0] != 15:
            raise ValueError(f"expected the 15 sample sites {SAMPLE_SITES} (or the first 14)")
        return x

    def evaluate(self, sites, diagnostics=False):
        """One model evaluation at the given sample-site values (dict or vector in SAMPLE_SITES order)."""
        x = self._site_vector(sites)
        theta = priors.theta_from_sites(x)
        r = self.like(theta)
        self.n_evals += 1
        nobs = r.nobs
        g_ll = priors.grad_sites_from_theta(r.dloglike, theta)
        g_mu = priors.grad_sites_from_theta(r.dlog_mu_sel, theta)
        mu_sel = math.exp(r.log_mu_sel) if math.isfinite(r.log_mu_sel) else float("nan")
        out = {
            "loglike": r.loglike, "selfactor": -nobs * r.log_mu_sel,                       # factors :383, :390
            "mbhmax": theta[7], "fpl": theta[9], "kappa": theta[12],                        # :288, :294, :301
            "log_mu_sel": r.log_mu_sel, "neff_sel": r.neff_sel, "neff": r.neff,             # :389, :394, :401
            "R": nobs / mu_sel + math.sqrt(nobs) / mu_sel * x[14],                          # :399
            "dloglike_dsite": g_ll, "dselfactor_dsite": -nobs * g_mu,
            "nobs": nobs, "theta": theta,
        }
        if diagnostics:
            out.update(self.diagnostics(theta, out["R"]))
        return out

    __call__ = evaluate

    # ---- potential energy in unconstrained space, as numpyro's NUTS sees the model
    def potential(self, u):
        """U(u) = -[log prior(x(u)) + log|dx/du| + loglike + selfactor] and dU/du (15-dim, incl. R_unit).

        This is the sampler's inner loop: one raw library call and plain-float host arithmetic.  The third return
        value is a light record; `deterministics(record)` expands it to the site names of `evaluate`."""
        x, dx, lpj, glp, dlj = priors.potential_terms(u)
        h, Om, w, a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp, r_unit = x
        fpl = math.exp(log_fpl)
        theta = (h, Om, w, a, b, c, mpisn, mpisn + dmbhmax, sigma, fpl, beta, lam, lam + dkappa, zp)
        out = self.like.raw(theta) if hasattr(self.like, "raw") else None
        if out is None:   # sharded torch-exchange path: no raw entry point
            ev = self.evaluate(np.array(x))
            logl = ev["loglike"] + ev["selfactor"]
            if not (math.isfinite(logl) and math.isfinite(lpj)):
                return math.inf, np.zeros(priors.NSITES), ev
            g = np.array(glp)
            g[:14] += ev["dloglike_dsite"] + ev["dselfactor_dsite"]
            return -(lpj + logl), -(g * np.array(dx) + np.array(dlj)), ev
        self.n_evals += 1
        nobs = out[36]
        loglike, log_mu = out[0], out[1]
        logl = loglike - nobs * log_mu
        rec = (theta, r_unit, out[:40].copy(), out[40:].copy())
        if not (math.isfinite(logl) and math.isfinite(lpj)):
            return math.inf, np.zeros(priors.NSITES), rec
        gt = out[4:18] - nobs * out[19:33]            # d(loglike + selfactor)/d theta
        g = [gt[0], gt[1], gt[2], gt[3], gt[4], gt[5], gt[6] + gt[7], gt[7], gt[8], gt[10], fpl * gt[9],
             gt[11] + gt[12], gt[12], gt[13], 0.0]    # chain rule to the sites (priors.grad_sites_from_theta)
        grad = np.array([-((g[i] + glp[i]) * dx[i] + dlj[i]) for i in range(15)])
        return -(lpj + logl), grad, rec

    @staticmethod
    def deterministics(rec):
        """Expand the light record of `potential` into the reference's deterministic / factor names."""
        if isinstance(rec, dict):
            return rec
        theta, r_unit, hdr, neff = rec
        nobs = hdr[36]
        mu = math.exp(hdr[1]) if math.isfinite(hdr[1]) else float("nan")
        return {"loglike": hdr[0], "selfactor": -nobs * hdr[1], "neff_sel": hdr[3], "neff": neff,
                "R": nobs / mu + math.sqrt(nobs) / mu * r_unit, "mbhmax": theta[7], "fpl": theta[9],
                "kappa": theta[12]}

    # ---- output-only curves (:403-406) from the tables the device built for this theta
    def diagnostics(self, theta, R):
        t = self._local.tables()
        sc = t["scalars"]
        log_norm, rate_log_norm, lpn = sc[18], sc[19], sc[6]
        h, Om, w, c, mbhmax, sigma, beta, lam, kappa, zp = (theta[0], theta[1], theta[2], theta[5], theta[7],
                                                             theta[8], theta[10], theta[11], theta[12], theta[13])
        mbh_grid = np.linspace(3.0, mbhmax + 7 * sigma, 256)
        G = t["log_dN_grid"]

        def log_dndm(m):   # LogDNDM.__call__ :140-151
            m = np.asarray(m, dtype=np.float64)
            with np.errstate(divide="ignore", over="ignore"):
                ld = np.interp(m, mbh_grid, G)
                ld = np.where((m <= mbh_grid[0]) | (m >= mbh_grid[-1]), -np.inf, ld)
                turn = math.log(2) - np.log1p(np.exp(-(m - mbhmax) / (0.05 * mbhmax)))
                ld = np.logaddexp(ld, -c * np.log(m / mbhmax) + lpn + turn)
                ld = np.where(m < 5.0, -np.inf, ld)
            return ld + log_norm

        def log_dndv(z):   # LogDNDV.__call__ :170-173
            z = np.asarray(z, dtype=np.float64)
            return lam * np.log1p(z) - np.log1p(((1 + z) / (1 + zp)) ** kappa) + rate_log_norm

        def log_dN(m1, q, z):   # LogDNDMDQDV.__call__ :202-210
            m1, q, z = np.broadcast_arrays(np.asarray(m1, float), np.asarray(q, float), np.asarray(z, float))
            m2 = q * m1
            with np.errstate(divide="ignore"):
                return (log_dndm(m1) + log_dndm(m2) + beta * np.log((m1 + m2) / (MREF * (1 + QREF))) + np.log(m1)
                        + log_dndv(z))

        zg = coords["z_grid"]
        opz = 1 + zg
        return {
            "mdNdmdVdt_fixed_qz": coords["m_grid"] * R * np.exp(log_dN(coords["m_grid"], QREF, ZREF)),
            "dNdqdVdt_fixed_mz": MREF * R * np.exp(log_dN(MREF, coords["q_grid"], ZREF)),
            "dNdVdt_fixed_mq": MREF * R * np.exp(log_dN(MREF, QREF, zg)),
            "hz": h * np.sqrt(Om * opz ** 3 + (1 - Om) * opz ** (3 * (1 + w))),      # :253-256
        }

    def close(self):
        self._local.close()


FIXED_SITES = ("a", "b", "c", "mpisn", "dmbhmax", "sigma", "beta", "log_fpl", "lam", "dkappa", "zp", "R_unit")


class PopModel:
    """The reference's fixed-cosmology model `pop_model` (intensity_models.py:313-355) bound to its data:
    source-frame (m1s, qs, zs, pdraw) and the theta-independent table `dVdzdt_interp` that the reference builds from
    astropy's Planck18 on zinterp = expm1(linspace(log1p(0), log1p(100), 1024)) (:323-325; astropy is not
    available here, so the caller supplies the 1024 numbers).  Sites: mass_parameters, redshift_parameters, R_unit."""

    def __init__(self, m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw, dVdzdt_interp, device=0):
        self.like = Hyperlikelihood(m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw, device=device,
                                    fixed_dvdzdt=dVdzdt_interp)

    def evaluate(self, sites):
        if isinstance(sites, dict):
            x = np.array([float(sites.get(k, 0.0)) for k in FIXED_SITES])
        else:
            x = np.asarray(sites, dtype=np.float64).ravel()
        full = np.concatenate([[0.7, 0.3, -1.0], x[:11]])     # (h, Om, w) are placeholders: ignored by the kernel
        theta = priors.theta_from_sites(full)
        r = self.like(theta)
        nobs = r.nobs
        mu_sel = math.exp(r.log_mu_sel) if math.isfinite(r.log_mu_sel) else float("nan")
        return {
            "loglike": r.loglike, "selfactor": -nobs * r.log_mu_sel, "mbhmax": theta[7], "fpl": theta[9],
            "kappa": theta[12], "log_mu_sel": r.log_mu_sel, "neff_sel": r.neff_sel, "neff": r.neff,
            "R": nobs / mu_sel + math.sqrt(nobs) / mu_sel * (x[11] if x.shape[0] > 11 else 0.0),
            "dloglike_dsite": priors.grad_sites_from_theta(r.dloglike, theta)[3:],
            "dselfactor_dsite": -nobs * priors.grad_sites_from_theta(r.dlog_mu_sel, theta)[3:],
            "nobs": nobs, "theta": theta,
        }

    __call__ = evaluate

    # ---- potential energy in unconstrained space (12 sites), as numpyro's NUTS sees pop_model
    dim = len(FIXED_SITES)

    @staticmethod
    def constrain(u):
        return priors.constrain(u, priors.FIXED_SITES)

    def potential(self, u):
        """U(u) = -[log prior + log|dx/du| + loglike + selfactor] and dU/du for the 12 sites of `pop_model`
        (a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp, R_unit); same structure as
        `PopCosmoModel.potential`."""
        x, dx, lpj, glp, dlj = priors.potential_terms(u, priors.FIXED_SITES)
        a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp, r_unit = x
        fpl = math.exp(log_fpl)
        theta = (0.7, 0.3, -1.0, a, b, c, mpisn, mpisn + dmbhmax, sigma, fpl, beta, lam, lam + dkappa, zp)
        out = self.like.raw(theta)
        self.n_evals = getattr(self, "n_evals", 0) + 1
        nobs = out[36]
        logl = out[0] - nobs * out[1]
        rec = (theta, r_unit, out[:40].copy(), out[40:].copy())
        if not (math.isfinite(logl) and math.isfinite(lpj)):
            return math.inf, np.zeros(self.dim), rec
        gt = out[4:18] - nobs * out[19:33]            # d(loglike + selfactor)/d theta; entries 0..2 are zero here
        g = [gt[3], gt[4], gt[5], gt[6] + gt[7], gt[7], gt[8], gt[10], fpl * gt[9], gt[11] + gt[12], gt[12], gt[13], 0.0]
        grad = np.array([-((g[i] + glp[i]) * dx[i] + dlj[i]) for i in range(self.dim)])
        return -(lpj + logl), grad, rec

    deterministics = staticmethod(lambda rec: PopCosmoModel.deterministics(rec))

    def close(self):
        self.like.close()


def pop_model(m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw, dVdzdt_interp=None, **kwargs):
    """Same positional signature as the reference's fixed-cosmology model (intensity_models.py:313)."""
    if dVdzdt_interp is None:
        raise ValueError("pop_model needs dVdzdt_interp (1024 values on the reference's zinterp grid): the reference "
                         "takes it from astropy's Planck18, which is not available here")
    return PopModel(m1s, qs, zs, pdraw, m1s_sel, qs_sel, zs_sel, pdraw_sel, Ndraw, dVdzdt_interp, **kwargs)


def pop_cosmo_model(m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, **kwargs):
    """Same positional signature as the reference model (intensity_models.py:357); returns the bound model."""
    return PopCosmoModel(m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, **kwargs)

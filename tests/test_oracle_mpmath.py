"""Third cross-check of the golden vectors: an independent 40-digit restatement of the forward
pass (oracle/mp_forward.py, mpmath) must reproduce the values minted from the reference source in fp64, and its
finite differences (exact at 40 digits) the reference's gradients with respect to all 14 sample sites.  CPU only."""
import os

import numpy as np
import pytest

mpmath = pytest.importorskip("mpmath")
from mpmath import mpf  # noqa: E402

from oracle import mp_forward as mpfw  # noqa: E402


def _load(golden_dir, name="tiny"):
    g = np.load(os.path.join(golden_dir, f"pop_cosmo_{name}.npz"))
    data = (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"],
            g["pdraw_sel"], float(g["Ndraw"]))
    return g, data


def _theta_from_sites(s):
    """(h, Om, w, a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp) -> kernel parameters
    (intensity_models.py:288, :294, :301)."""
    h, Om, w, a, b, c, mpisn, dmbhmax, sigma, beta, log_fpl, lam, dkappa, zp = s
    return (h, Om, w, a, b, c, mpisn, mpisn + dmbhmax, sigma, mpmath.exp(log_fpl), beta, lam, lam + dkappa, zp)


def _sites_from_theta(th):
    h, Om, w, a, b, c, mpisn, mbhmax, sigma, fpl, beta, lam, kappa, zp = (mpf(float(t)) for t in th)
    return [h, Om, w, a, b, c, mpisn, mbhmax - mpisn, sigma, beta, mpmath.log(fpl), lam, kappa - lam, zp]


@pytest.mark.parametrize("k", (0, 3))
def test_goldens_match_40_digit_forward(golden_dir, k):
    g, data = _load(golden_dir)
    r = mpfw.forward(g["thetas"][k], data)
    for key in ("loglike", "log_mu_sel", "neff_sel"):
        ref = float(g["ref_" + key][k])
        assert abs(float(r[key]) - ref) <= 1e-12 * max(1.0, abs(ref)), (key, float(r[key]), ref)
    assert np.allclose([float(v) for v in r["neff"]], g["ref_neff"][k], rtol=1e-12, atol=0)


def test_reference_gradients_match_40_digit_finite_differences(golden_dir):
    """Central differences at a relative step of 1e-15 in 40-digit arithmetic: truncation ~1e-30, rounding ~1e-25.
    All 14 sites; the five that move the PISN table rebuild it (5 s each side), the other nine reuse it."""
    g, data = _load(golden_dir)
    k = 0
    sites = _sites_from_theta(g["thetas"][k])
    base = mpfw.forward(_theta_from_sites(sites), data)
    grid = base["pisn_grid"]
    pisn_sites = {3, 4, 6, 7, 8}          # a, b, mpisn, dmbhmax, sigma
    d_ll, d_mu = [], []
    for i in range(14):
        step = mpf("1e-15") * max(abs(sites[i]), mpf(1))
        vals = []
        for sgn in (1, -1):
            s = list(sites)
            s[i] = s[i] + sgn * step
            vals.append(mpfw.forward(_theta_from_sites(s), data, pisn_grid=None if i in pisn_sites else grid))
        d_ll.append(float((vals[0]["loglike"] - vals[1]["loglike"]) / (2 * step)))
        d_mu.append(float((vals[0]["log_mu_sel"] - vals[1]["log_mu_sel"]) / (2 * step)))
    ref_ll, ref_mu = g["ref_dloglike_dsite"][k], g["ref_dlog_mu_sel_dsite"][k]
    scale = max(1.0, float(np.max(np.abs(ref_ll))))
    assert np.all(np.abs(np.array(d_ll) - ref_ll) <= 1e-10 * np.maximum(np.abs(ref_ll), scale)), (d_ll, ref_ll)
    assert np.all(np.abs(np.array(d_mu) - ref_mu) <= 1e-10 * np.maximum(np.abs(ref_mu), 1.0)), (d_mu, ref_mu)

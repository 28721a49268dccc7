"""The restated oracle (oracle/bump_oracle.py) against the golden vectors minted from the UNMODIFIED
reference source (tests/golden/make_golden.py -> oracle/run_reference.py).  CPU only."""
import math
import os

import numpy as np
import pytest

from oracle import bump_oracle as bo

CASES = ("tiny", "small")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"pop_cosmo_{name}.npz"))


def _data(g):
    return (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"],
            g["pdraw_sel"], float(g["Ndraw"]))


def _close(a, b, rtol=1e-11, floor=1.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_goldens(golden_dir, name):
    g = _load(golden_dir, name)
    data = _data(g)
    for k, th in enumerate(g["thetas"]):
        r = bo.evaluate(th, data, grad=True)
        assert _close(r["loglike"], g["ref_loglike"][k]), (k, r["loglike"], g["ref_loglike"][k])
        assert _close(r["log_mu_sel"], g["ref_log_mu_sel"][k])
        assert _close(r["selfactor"], g["ref_selfactor"][k])
        assert _close(r["neff_sel"], g["ref_neff_sel"][k], rtol=1e-10)
        assert _close(r["neff"], g["ref_neff"][k], rtol=1e-10)
        gs = bo.grad_sites_from_theta(r["dloglike"], th)
        gm = bo.grad_sites_from_theta(r["dlog_mu_sel"], th)
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(gs, g["ref_dloglike_dsite"][k], rtol=1e-10, floor=scale)
        assert _close(gm, g["ref_dlog_mu_sel_dsite"][k], rtol=1e-10, floor=1.0)


@pytest.mark.parametrize("name", CASES)
def test_oracle_tables_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    for k, th in enumerate(g["thetas"]):
        t = bo.tables(th)
        for key in ("zinterp", "dcinterp", "dlinterp", "ddlinterp", "dvcinterp", "mbh_grid", "log_dN_grid"):
            assert _close(t[key], g["tab_" + key][k], rtol=1e-12, floor=1e-3), key


def test_identities():
    """Internal consistency properties of the reference model."""
    import torch
    th = torch.tensor([0.7, 0.3, -1.0, 1.8, -0.71, 2.9, 31.0, 36.0, 2.3, 0.21, -2.2, 4.7, 7.0, 3.0], dtype=bo.F)
    cosmo, log_dN = bo.build_model(th)
    mf = log_dN.log_dndm
    # m dN/dm = 1 at m = mref after normalisation (intensity_models.py:138)
    assert abs(float(mf(torch.tensor(30.0, dtype=bo.F))) + math.log(30.0)) < 1e-12
    # log_dndv(zref) = 0 (:168)
    assert abs(float(log_dN.log_dndv(torch.tensor(0.0, dtype=bo.F)))) < 1e-14
    # log_smooth_turnon(mmin, mmin) = 0 (:54)
    assert abs(float(bo.log_smooth_turnon(torch.tensor(36.0, dtype=bo.F), torch.tensor(36.0, dtype=bo.F)))) < 1e-15
    # mean_mbh_from_mco continuous at mco = mpisn (:22-25)
    mp, mb = torch.tensor(31.0, dtype=bo.F), torch.tensor(36.0, dtype=bo.F)
    lo = bo.mean_mbh_from_mco(torch.tensor(31.0 - 1e-9, dtype=bo.F), mp, mb)
    hi = bo.mean_mbh_from_mco(torch.tensor(31.0, dtype=bo.F), mp, mb)
    assert abs(float(lo - hi)) < 1e-8
    # dl table: dl_0 = 0, ddl_0 = dH
    assert float(cosmo.dlinterp[0]) == 0.0
    assert abs(float(cosmo.ddlinterp[0]) - 2.99792 / 0.7) < 1e-13


def test_wa_zero_reduces_to_reference(golden_dir):
    g = _load(golden_dir, "tiny")
    data = _data(g)
    th = g["thetas"][0]
    r0 = bo.evaluate(th, data, grad=True)
    r1 = bo.evaluate(th, data, grad=True, wa=0.0)
    assert _close(r1["loglike"], r0["loglike"], rtol=1e-13)
    assert _close(r1["dloglike"], r0["dloglike"], rtol=1e-12)
    assert np.isfinite(r1["dloglike_dwa"]) and np.isfinite(r1["dlog_mu_sel_dwa"])


def test_gradient_vs_finite_differences(golden_dir):
    """Mass / rate parameters only: the likelihood is merely C0 in (h, Om, w) (knot crossings)."""
    g = _load(golden_dir, "tiny")
    data = _data(g)
    th = np.array(g["thetas"][0])
    r = bo.evaluate(th, data, grad=True)
    for i in (3, 4, 5, 6, 8, 9, 10, 11, 12, 13):
        eps = 1e-6 * max(1.0, abs(th[i]))
        tp, tm = th.copy(), th.copy()
        tp[i] += eps
        tm[i] -= eps
        fd = (bo.evaluate(tp, data, grad=False)["logl"] - bo.evaluate(tm, data, grad=False)["logl"]) / (2 * eps)
        assert abs(fd - r["dlogl"][i]) <= 2e-6 * max(1.0, abs(fd)), (i, fd, r["dlogl"][i])


def test_fixed_cosmology_oracle_matches_reference_pop_model(golden_dir):
    """oracle.evaluate_fixed against the golden minted from the unmodified reference `pop_model`
    (intensity_models.py:313-355)."""
    g = np.load(os.path.join(golden_dir, "pop_fixed_small.npz"))
    data = (g["m1s"], g["qs"], g["zs"], g["pdraw"], g["m1s_sel"], g["qs_sel"], g["zs_sel"], g["pdraw_sel"],
            float(g["Ndraw"]))
    for k, th in enumerate(g["thetas"]):
        r = bo.evaluate_fixed(th, data, g["dvdzdt_interp"])
        assert _close(r["loglike"], g["ref_loglike"][k]) and _close(r["log_mu_sel"], g["ref_log_mu_sel"][k])
        assert _close(r["neff_sel"], g["ref_neff_sel"][k], rtol=1e-10) and _close(r["neff"], g["ref_neff"][k], rtol=1e-10)
        gs = bo.grad_sites_from_theta(r["dloglike"], th)[3:]
        gm = bo.grad_sites_from_theta(r["dlog_mu_sel"], th)[3:]
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(gs, g["ref_dloglike_dsite"][k], rtol=1e-10, floor=scale)
        assert _close(gm, g["ref_dlog_mu_sel_dsite"][k], rtol=1e-10)
        assert np.all(r["dloglike"][:3] == 0) and np.all(r["dlog_mu_sel"][:3] == 0)

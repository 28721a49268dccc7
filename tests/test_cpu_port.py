"""The fused C++/OpenMP CPU port (oracle/bump_cpu.cpp: bench.py's CPU baseline) against the goldens minted from the
unmodified reference and against the torch oracle.  A third, independently written evaluation of the path."""
import os

import numpy as np
import pytest

from oracle import bump_cpu, bump_oracle as bo


def _close(a, b, rtol=1e-10, floor=1.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), floor)))


@pytest.mark.parametrize("name", ("tiny", "small"))
def test_cpu_port_matches_reference_goldens(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"pop_cosmo_{name}.npz"))
    data = (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"], g["pdraw_sel"],
            float(g["Ndraw"]))
    port = bump_cpu.CpuPort(*data)
    for k, th in enumerate(g["thetas"]):
        r = port.evaluate(th)
        assert _close(r["loglike"], g["ref_loglike"][k]) and _close(r["log_mu_sel"], g["ref_log_mu_sel"][k])
        assert _close(r["neff_sel"], g["ref_neff_sel"][k]) and _close(r["neff"], g["ref_neff"][k])
        scale = max(1.0, float(np.max(np.abs(g["ref_dloglike_dsite"][k]))))
        assert _close(bo.grad_sites_from_theta(r["dloglike"], th), g["ref_dloglike_dsite"][k], floor=scale), k
        assert _close(bo.grad_sites_from_theta(r["dlog_mu_sel"], th), g["ref_dlog_mu_sel_dsite"][k]), k
    port.close()


def test_cpu_port_matches_oracle_on_ragged_catalog_and_thread_counts():
    from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
    cat = make_catalog("gwtc3", nobs=6, nsamp=777, nsel=5001, seed=41)
    port = bump_cpu.CpuPort(*cat.as_args())
    o = bo.evaluate(THETA_DEFAULT, cat.as_args(), grad=True)
    for nt in (1, 3, 0):
        r = port.evaluate(THETA_DEFAULT, nthreads=nt)
        assert _close(r["loglike"], o["loglike"]) and _close(r["log_mu_sel"], o["log_mu_sel"])
        assert _close(r["log_mu2"], o["log_mu2"]) and _close(r["neff"], o["neff"])
        assert _close(r["dloglike"], o["dloglike"], floor=max(1.0, float(np.max(np.abs(o["dloglike"])))))
        assert _close(r["dlog_mu_sel"], o["dlog_mu_sel"])
    port.close()

"""The library's C++ NUTS driver (`bump_nuts_chain`) bound to the CUDA hot path: its potential must be the host
mirror's potential, and a short run must behave like the Python driver's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from bumpcosmology_b200 import intensity_models as im
    from bumpcosmology_b200.catalogs import make_catalog
    cat = make_catalog("gwtc3_nuts", nobs=24, nsamp=1024, nsel=20000)
    m = im.pop_cosmo_model(*cat.as_args())
    yield m
    m.close()


def test_native_potential_equals_host_mirror_potential(model):
    """Every recorded draw carries the potential the C++ driver computed there: recompute it with the Python
    potential (priors.py transforms + the same library evaluation) at the same unconstrained point."""
    from bumpcosmology_b200 import nuts, priors
    c = nuts.run_chain_native(model, num_warmup=60, num_samples=40, seed=7)
    assert np.all(np.isfinite(c["u"])) and np.all(np.isfinite(c["stats"]["potential"]))
    for j in range(0, 40, 5):
        U, g, rec = model.potential(c["u"][j])
        assert abs(U - c["stats"]["potential"][j]) <= 1e-11 * max(1.0, abs(U))
        det = model.deterministics(rec)
        for name in ("loglike", "selfactor", "neff_sel", "R", "mbhmax", "fpl", "kappa"):
            assert abs(det[name] - c["deterministic"][name][j]) <= 1e-11 * max(1.0, abs(det[name])), name
        assert abs(np.min(det["neff"]) - c["deterministic"]["neff_min"][j]) <= 1e-11 * np.min(det["neff"])
        x = priors.constrain(c["u"][j])[0]
        assert np.allclose(x, c["x"][j], rtol=1e-13, atol=0)


def test_native_and_python_chains_agree_statistically(model):
    from bumpcosmology_b200 import nuts
    a = nuts.run_chain_native(model, num_warmup=300, num_samples=300, seed=21)
    b = nuts.run_chain(model, num_warmup=300, num_samples=300, seed=21)
    assert a["stats"]["diverging"].sum() <= 3 and b["stats"]["diverging"].sum() <= 3
    assert 0.6 < a["stats"]["accept"].mean() < 0.97
    assert abs(a["stats"]["depth"].mean() - b["stats"]["depth"].mean()) < 1.0
    sd = b["x"].std(0)
    # two independent chains of ~300 correlated draws each: means within a generous multiple of the spread
    assert np.all(np.abs(a["x"].mean(0) - b["x"].mean(0)) < 0.6 * sd)


def test_fit_driver_end_to_end_from_tables(tmp_path):
    """The reference's run_cosmo_fit.py flow on mock tables: tables -> detector frame -> 2 chains of the C++ driver ->
    trace file with the reference's site and deterministic names."""
    import pandas as pd

    from bumpcosmology_b200 import priors, run_cosmo_fit
    from mock_tables import make_tables
    pe, sel = make_tables()
    pd.DataFrame(pe).to_parquet(tmp_path / "pe-samples.parquet")
    pd.DataFrame(sel).to_parquet(tmp_path / "selection-samples.parquet")
    out = tmp_path / "trace_cosmo.npz"
    trace = run_cosmo_fit.main(["--pe", str(tmp_path / "pe-samples.parquet"), "--sel",
                                str(tmp_path / "selection-samples.parquet"), "--out", str(out), "--nmcmc", "120",
                                "--nchain", "2"])
    z = np.load(out)
    assert list(z["site_names"]) == list(priors.SITE_NAMES)
    assert z["posterior"].shape == (2, 120, 15) and np.all(np.isfinite(z["posterior"]))
    for k in ("det_loglike", "det_selfactor", "det_neff_sel", "det_R", "det_mbhmax", "det_fpl", "det_kappa",
              "stat_accept", "stat_depth", "stat_diverging"):
        assert z[k].shape == (2, 120), k
    x = z["posterior"]
    names = list(z["site_names"])
    assert np.allclose(z["det_mbhmax"], x[:, :, names.index("mpisn")] + x[:, :, names.index("dmbhmax")])
    assert np.allclose(z["det_fpl"], np.exp(x[:, :, names.index("log_fpl")]))
    assert 0.5 < trace["stat_accept"].mean() <= 1.0

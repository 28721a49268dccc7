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
    # the trace is a drop-in for the reference's trace_cosmo.nc: every sample site and every numpyro.deterministic of
    # pop_cosmo_model (intensity_models.py:282-309,288,294,301,394,398-406) under the reference's name and shape
    post = run_cosmo_fit.posterior_variables(trace)
    reference_names = set(priors.SITE_NAMES) | set(run_cosmo_fit.REFERENCE_DETERMINISTICS)
    assert reference_names <= set(post), reference_names - set(post)
    nobs = int(trace["nobs"])
    assert post["neff"].shape == (2, 120, nobs) and np.all(post["neff"] > 0)
    for k in ("mdNdmdVdt_fixed_qz", "dNdqdVdt_fixed_mz", "dNdVdt_fixed_mq", "hz"):
        assert post[k].shape == (2, 120, 128) and np.all(np.isfinite(post[k])), k
    assert np.allclose(post["neff"].min(axis=2), trace["det_neff_min"], rtol=1e-12)   # the same draws, re-evaluated
    h, om, w = (x[:, :, names.index(k)] for k in ("h", "Om", "w"))
    from bumpcosmology_b200.intensity_models import coords
    opz = 1 + coords["z_grid"]
    hz = h[..., None] * np.sqrt(om[..., None] * opz ** 3 + (1 - om[..., None]) * opz ** (3 * (1 + w[..., None])))
    assert np.allclose(post["hz"], hz, rtol=1e-12)                                          # h E(z) (:253-256, :406)
    # dNdVdt_fixed_mq at z = 0 is mref * R * exp(log_dN(mref, qref, zref)) = mref * R * exp(0 + ... ) (:404-405)
    assert np.all(post["dNdVdt_fixed_mq"][:, :, 0] > 0)


def test_fixed_cosmology_potential_gradient_and_fit_driver(tmp_path):
    """pop_model (the reference's run_fit.py path): the 12-site potential's gradient equals central differences, and the
    fit driver runs end to end from tables with the Python NUTS driver."""
    import pandas as pd

    from bumpcosmology_b200 import inputs, intensity_models as im, run_fit
    from mock_tables import make_tables
    pe, sel = make_tables(nobs=8, nsamp=128, nsel=5000)
    _, (m1s, qs, zs, wts) = inputs.group_events(pe["evt"], pe["m1"], pe["q"], pe["z"], pe["wt"])
    model = im.pop_model(m1s, qs, zs, wts, sel["m1"], sel["q"], sel["z"], sel["pdraw"], float(sel["ndraw"][0]),
                         dVdzdt_interp=inputs.FlatLCDM().dVdzdt_interp())
    rng = np.random.default_rng(2)
    u = rng.uniform(-0.5, 0.5, 12)
    U, g, rec = model.potential(u)
    assert np.isfinite(U) and g.shape == (12,)
    for i in range(12):
        up, um = u.copy(), u.copy()
        up[i] += 1e-7
        um[i] -= 1e-7
        fd = (model.potential(up)[0] - model.potential(um)[0]) / 2e-7
        # mpisn, dmbhmax and sigma move the knots of the PISN table: a sample crossing a knot between the two
        # evaluations adds a slope jump to the difference quotient (the analytic gradient is pinned to the reference at
        # 1e-10 in test_gpu_parity.py; this test is about the chain rule of the 12-site potential)
        tol = 5e-3 if i in (3, 4, 5) else 2e-5
        assert abs(fd - g[i]) <= tol * max(1.0, abs(g[i])), (i, fd, g[i])
    det = model.deterministics(rec)
    x = model.constrain(u)[0]
    assert abs(det["mbhmax"] - (x[3] + x[4])) < 1e-12
    model.close()
    pd.DataFrame(pe).to_parquet(tmp_path / "pe.parquet")
    pd.DataFrame(sel).to_parquet(tmp_path / "sel.parquet")
    trace = run_fit.main(["--pe", str(tmp_path / "pe.parquet"), "--sel", str(tmp_path / "sel.parquet"), "--out",
                          str(tmp_path / "trace.npz"), "--nmcmc", "60", "--nchain", "2"])
    assert trace["posterior"].shape == (2, 60, 12) and np.all(np.isfinite(trace["posterior"]))
    assert list(trace["site_names"]) == list(im.FIXED_SITES)
    assert 0.05 < trace["stat_accept"].mean() <= 1.0   # 60 warm-up steps: barely adapted

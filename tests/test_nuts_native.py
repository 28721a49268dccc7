"""The C++ NUTS driver of the library (csrc/bump_nuts.cpp, `bump_nuts_chain_cb`) on CPU: a correlated Gaussian stands
in for the model (the callback entry point needs no GPU), and the Python driver is the cross-check."""
import numpy as np

from bumpcosmology_b200 import nuts


def _gaussian(dim=15, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((dim, dim))
    cov = a @ a.T / dim + 0.5 * np.eye(dim)
    prec = np.linalg.inv(cov)
    mu = rng.standard_normal(dim)

    def potential(u):
        d = u - mu
        g = prec @ d
        return 0.5 * float(d @ g), g

    return mu, cov, potential


def test_native_nuts_recovers_gaussian_moments_and_adapts_the_mass_matrix():
    mu, cov, f = _gaussian()
    chains = [nuts.run_chain_native_fn(f, 15, num_warmup=300, num_samples=500, seed=100 + c) for c in range(2)]
    u = np.concatenate([c["u"] for c in chains])
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(u.mean(0) - mu) < 0.25 * sd)
    assert np.all(np.abs(u.std(0) / sd - 1) < 0.2)
    for c in chains:
        assert 0.6 < c["stats"]["accept"].mean() < 0.97
        assert c["stats"]["diverging"].sum() == 0
        assert 1 <= c["stats"]["depth"].mean() <= 5
        # the dense inverse mass matrix is an estimate of the posterior covariance
        assert np.all(np.abs(np.diag(c["inverse_mass"]) / np.diag(cov) - 1) < 0.5)
        assert c["n_leapfrog_total"] == c["n_evals"] - 1 or c["n_leapfrog_total"] <= c["n_evals"]
    x = np.stack([c["u"] for c in chains])
    ess = np.array([nuts.ess_bulk(x[:, :, i]) for i in range(15)])
    rhat = np.array([nuts.split_rhat(x[:, :, i]) for i in range(15)])
    assert np.all(ess > 200) and np.all(rhat < 1.05)


def test_native_and_python_drivers_agree_statistically():
    mu, cov, f = _gaussian(dim=6, seed=3)

    class M:
        def potential(self, u):
            U, g = f(u)
            ev = dict(loglike=0.0, selfactor=0.0, neff_sel=1.0, R=1.0, mbhmax=1.0, fpl=1.0, kappa=1.0, neff=np.ones(1))
            return U, g, ev

    import bumpcosmology_b200.priors as priors
    old = priors.NSITES
    try:   # the Python driver takes its dimension from the model's site table
        priors.NSITES = 6
        priors_constrain = priors.constrain
        priors.constrain = lambda u: (np.asarray(u), None, 0.0, None)
        py = nuts.run_chain(M(), num_warmup=300, num_samples=600, seed=5)
    finally:
        priors.NSITES = old
        priors.constrain = priors_constrain
    cc = nuts.run_chain_native_fn(f, 6, num_warmup=300, num_samples=600, seed=5)
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(py["u"].mean(0) - cc["u"].mean(0)) < 0.3 * sd)
    assert np.all(np.abs(py["u"].std(0) / cc["u"].std(0) - 1) < 0.25)
    # same adaptation targets: step sizes and tree depths of the two drivers are comparable
    assert 0.5 < py["step_size"] / cc["step_size"] < 2.0
    assert abs(py["stats"]["depth"].mean() - cc["stats"]["depth"].mean()) < 1.0


def test_native_nuts_divergent_potential_is_flagged_not_fatal():
    def funnel(u):   # non-finite outside |u| < 3: transitions that leave the region count as divergent
        if np.any(np.abs(u) > 3.0):
            return float("inf"), np.zeros_like(u)
        return 0.5 * float(u @ u), u.copy()

    c = nuts.run_chain_native_fn(funnel, 2, num_warmup=100, num_samples=200, seed=1, init=np.zeros(2))
    assert np.all(np.abs(c["u"]) <= 3.0)
    assert np.all(np.isfinite(c["stats"]["potential"]))


def test_native_nuts_is_reproducible_and_validates_arguments():
    mu, cov, f = _gaussian(dim=4, seed=9)
    a = nuts.run_chain_native_fn(f, 4, num_warmup=80, num_samples=60, seed=42)
    b = nuts.run_chain_native_fn(f, 4, num_warmup=80, num_samples=60, seed=42)
    c = nuts.run_chain_native_fn(f, 4, num_warmup=80, num_samples=60, seed=43)
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["stats"]["depth"], b["stats"]["depth"])
    assert not np.array_equal(a["u"], c["u"])
    from bumpcosmology_b200._lib import BumpError
    import pytest
    with pytest.raises(BumpError, match="dimension"):
        nuts.run_chain_native_fn(lambda u: (0.0, np.zeros(40)), 40, num_warmup=5, num_samples=5)
    with pytest.raises(BumpError, match="finite starting point"):
        nuts.run_chain_native_fn(lambda u: (float("nan"), np.zeros(2)), 2, num_warmup=5, num_samples=5)


def test_native_nuts_diagonal_mass_matrix_option():
    mu, cov, f = _gaussian(dim=5, seed=2)
    c = nuts.run_chain_native_fn(f, 5, num_warmup=200, num_samples=200, seed=3, dense_mass=False)
    m = c["inverse_mass"]
    assert np.allclose(m, np.diag(np.diag(m)))
    assert np.all(np.abs(c["u"].mean(0) - mu) < 0.4 * np.sqrt(np.diag(cov)))

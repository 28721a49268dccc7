"""Host NUTS driver and priors on CPU: a correlated Gaussian stands in for the model (no GPU needed)."""
import math

import numpy as np

from bumpcosmology_b200 import nuts, priors


class GaussianModel:
    """potential(u) of N(mu, Sigma) in 15 dimensions; carries the keys run_chain records."""

    def __init__(self, seed=0):
        rng = np.random.default_rng(seed)
        a = rng.standard_normal((15, 15))
        self.cov = a @ a.T / 15 + 0.5 * np.eye(15)
        self.prec = np.linalg.inv(self.cov)
        self.mu = rng.standard_normal(15)

    def potential(self, u):
        d = u - self.mu
        g = self.prec @ d
        ev = dict(loglike=0.0, selfactor=0.0, neff_sel=1.0, R=1.0, mbhmax=1.0, fpl=1.0, kappa=1.0, neff=np.ones(1))
        return 0.5 * float(d @ g), g, ev


def test_adaptation_windows_match_stan_schedule():
    start, ends = nuts.adaptation_windows(1000)
    assert start == 75 and ends == [100, 150, 250, 450, 950]


def test_nuts_recovers_gaussian_moments_and_ess():
    m = GaussianModel()
    r = nuts.run_mcmc(m, num_warmup=300, num_samples=400, num_chains=2, seed=11)
    u = np.concatenate([c["u"] for c in r["chains"]])
    sd = np.sqrt(np.diag(m.cov))
    assert np.all(np.abs(u.mean(0) - m.mu) < 0.25 * sd)
    assert np.all(np.abs(u.std(0) / sd - 1) < 0.2)
    acc = np.mean([c["stats"]["accept"].mean() for c in r["chains"]])
    assert 0.6 < acc < 0.97
    assert all(c["stats"]["diverging"].sum() == 0 for c in r["chains"])
    assert np.all(r["ess_bulk"] > 150) and np.all(r["rhat"] < 1.05)


def test_ess_of_iid_and_correlated_draws():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 1000))
    assert 3000 < nuts.ess_bulk(x) < 5000
    y = np.empty((4, 1000))
    y[:, 0] = x[:, 0]
    for t in range(1, 1000):     # AR(1), rho = 0.9 -> ESS ~ N (1-rho)/(1+rho)
        y[:, t] = 0.9 * y[:, t - 1] + math.sqrt(1 - 0.81) * x[:, t]
    assert 100 < nuts.ess_bulk(y) < 450


def test_prior_transforms_roundtrip_and_densities():
    from scipy import stats
    rng = np.random.default_rng(5)
    x = priors.sample_prior(rng)
    u = priors.unconstrain(x)
    x2, dx, lj, dlj = priors.constrain(u)
    assert np.allclose(x, x2, rtol=1e-12, atol=1e-12)
    # jacobian terms by finite differences
    for i in range(priors.NSITES):
        up, um = u.copy(), u.copy()
        up[i] += 1e-6
        um[i] -= 1e-6
        assert abs((priors.constrain(up)[0][i] - priors.constrain(um)[0][i]) / 2e-6 - dx[i]) < 1e-6 * max(1, abs(dx[i]))
        assert abs((priors.constrain(up)[2] - priors.constrain(um)[2]) / 2e-6 - dlj[i]) < 1e-6
    # truncated normal with the Phi normaliser (numpyro TruncatedNormal semantics)
    s = priors.SITES[0]
    ref = stats.truncnorm.logpdf(0.8, (0.35 - 0.7) / 0.2, (1.4 - 0.7) / 0.2, loc=0.7, scale=0.2)
    assert abs(s.log_prob(0.8) - ref) < 1e-12
    s = priors.SITES[8]   # sigma: one-sided
    ref = stats.truncnorm.logpdf(3.0, (1 - 2) / 2, np.inf, loc=2, scale=2)
    assert abs(s.log_prob(3.0) - ref) < 1e-12
    assert priors.SITES[10].log_prob(0.0) == -math.inf       # log_fpl outside its uniform support


def test_site_chain_rule_matches_oracle_helper():
    from oracle import bump_oracle as bo
    rng = np.random.default_rng(1)
    x = priors.sample_prior(rng)
    th = priors.theta_from_sites(x)
    s = dict(zip(priors.SITE_NAMES, x))
    assert np.allclose(th, bo.theta_from_sites(s))
    g = rng.standard_normal(14)
    assert np.allclose(priors.grad_sites_from_theta(g, th), bo.grad_sites_from_theta(g, th))


def test_fused_potential_terms_equal_constrain_plus_log_prior():
    rng = np.random.default_rng(9)
    for _ in range(100):
        u = rng.uniform(-4, 4, priors.NSITES)
        x, dx, lj, dlj = priors.constrain(u)
        lp, g = priors.log_prior(x)
        xs, dxs, tot, glp, dljs = priors.potential_terms(u)
        assert np.allclose(x, xs, rtol=1e-14) and np.allclose(dx, dxs, rtol=1e-13)
        assert abs(tot - (lp + lj)) < 1e-11 and np.allclose(g, glp, atol=1e-13) and np.allclose(dlj, dljs, atol=1e-14)

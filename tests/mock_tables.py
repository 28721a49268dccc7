"""Small source-frame tables in the reference's column schema (draw_pe_samples.py:24, draw_selection_samples.py:15)
for the driver tests: events well above the mass cut, Gaussian-ish posteriors, weights = 1."""
import numpy as np

from bumpcosmology_b200 import inputs


def make_tables(nobs=16, nsamp=256, nsel=20000, seed=3):
    rng = np.random.default_rng(seed)
    z0 = rng.uniform(0.1, 0.8, nobs)
    m10 = rng.uniform(15.0, 60.0, nobs)
    q0 = rng.uniform(0.6, 1.0, nobs)
    m1 = m10[:, None] * np.exp(0.08 * rng.standard_normal((nobs, nsamp)))
    q = np.clip(q0[:, None] + 0.08 * rng.standard_normal((nobs, nsamp)), 0.55, 1.0)
    z = z0[:, None] * np.exp(0.15 * rng.standard_normal((nobs, nsamp)))
    pe = {"m1": m1.ravel(), "q": q.ravel(), "z": z.ravel(), "wt": np.ones(nobs * nsamp),
          "evt": np.repeat(np.arange(nobs), nsamp)}
    zs = rng.uniform(0.02, 1.2, nsel)
    m1s = 12.0 * (1 - rng.uniform(size=nsel)) ** (-1 / 1.35)       # power law above 12
    m1s = np.minimum(m1s, 150.0)
    qs = rng.uniform(0.6, 1.0, nsel)
    sel = {"m1": m1s, "q": qs, "z": zs, "pdraw": 1.35 * 12.0 ** 1.35 * m1s ** -2.35 / 0.4 / 1.18,
           "ndraw": np.full(nsel, 10 * nsel)}
    return pe, sel


def direct_arguments(pe, sel):
    return inputs.model_arguments(pe, sel)

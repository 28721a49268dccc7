"""Synthetic catalogs: shapes, determinism, validity (every event keeps finite-weight samples), mass_floor."""
import numpy as np

from bumpcosmology_b200 import catalogs


def test_shapes_and_determinism():
    a = catalogs.make_catalog("small")
    b = catalogs.make_catalog("small")
    assert a.m1s_det.shape == (12, 256) and a.m1s_det_sel.shape == (4096,) and a.Ndraw == 40960.0
    for x, y in zip(a.as_args()[:8], b.as_args()[:8]):
        assert np.array_equal(x, y)
    assert catalogs.SHAPES["o5"][:3] == (5000, 10000, 10_000_000)
    assert catalogs.SHAPES["gwtc3"][:3] == (69, 4096, 200_000) and catalogs.SHAPES["o4"][:3] == (300, 8192, 1_000_000)


def test_inputs_are_positive_and_events_keep_weight():
    cat = catalogs.make_catalog("small")
    for x in cat.as_args()[:8]:
        assert np.all(np.isfinite(x)) and np.all(x > 0)
    assert np.all((cat.qs > 0) & (cat.qs <= 1)) and np.all((cat.qs_sel > 0) & (cat.qs_sel <= 1))
    cosmo = catalogs._FiducialCosmology()
    z = np.interp(cat.dls, cosmo.dl, cosmo.z)
    m2 = cat.qs * cat.m1s_det / (1 + z)
    assert np.all(np.median(m2, axis=1) >= 5.0)        # the reference's event selection, weighting.py:88-89


def test_mass_floor_keeps_every_sample_above_the_cut():
    cat = catalogs.make_catalog("small", mass_floor=8.0)
    cosmo = catalogs._FiducialCosmology()
    z = np.interp(cat.dls, cosmo.dl, cosmo.z)
    assert np.all(cat.qs * cat.m1s_det / (1 + z) >= 8.0 - 1e-9)
    zs = np.interp(cat.dls_sel, cosmo.dl, cosmo.z)
    assert np.all(cat.qs_sel * cat.m1s_det_sel / (1 + zs) >= 8.0 - 1e-9)
    assert catalogs.MASS_FLOOR["gwtc3_nuts"] == 8.0


def test_prior_draws_respect_support():
    th = catalogs.draw_prior_thetas(50, seed=1)
    assert np.all((th[:, 0] >= 0.35) & (th[:, 0] <= 1.4)) and np.all((th[:, 1] > 0) & (th[:, 1] < 1))
    assert np.all(th[:, 7] - th[:, 6] >= 0.5) and np.all(th[:, 8] >= 1.0) and np.all(th[:, 12] > th[:, 11])

"""Input-side helpers (host): detector-frame conversion, Jacobian, event grouping."""
import numpy as np
import pytest

from bumpcosmology_b200 import catalogs, inputs


def test_jacobian_matches_finite_differences():
    c = inputs.FlatLCDM()
    z = np.array([0.1, 0.5, 1.2, 2.5])
    m1 = np.array([30.0, 20.0, 45.0, 10.0])
    # d(m1_det, d_L)/d(m1, z) is triangular: det = (1+z) * d d_L/dz
    eps = 1e-4   # the distance table is piecewise linear: differences are good to ~1e-6
    ddl = (c.luminosity_distance(z + eps) - c.luminosity_distance(z - eps)) / (2 * eps)
    assert np.allclose(inputs.dm1sqz_dm1ddqdl(m1, 0.8, z, c), 1.0 / ((1 + z) * ddl), rtol=5e-6)


def test_fiducial_cosmology_agrees_with_the_catalog_generator():
    c = inputs.FlatLCDM(H0=67.66, Om0=0.30966)
    f = catalogs._FiducialCosmology()
    z = np.linspace(0.01, 3.0, 50)
    # catalogs.py uses the reference's c/H100 = 2.99792 constant; astropy's is 2.99792458
    assert np.allclose(c.luminosity_distance(z), np.interp(z, f.z, f.dl), rtol=2e-5)   # 4096-point table interpolation


def test_model_arguments_roundtrip():
    rng = np.random.default_rng(0)
    nobs, nsamp = 4, 16
    evt = np.repeat(np.array(["GW3", "GW1", "GW4", "GW2"]), nsamp)
    perm = rng.permutation(evt.size)
    pe = {"m1": rng.uniform(10, 50, evt.size), "q": rng.uniform(0.3, 1, evt.size), "z": rng.uniform(0.05, 1, evt.size),
          "wt": rng.uniform(0.5, 2, evt.size), "evt": evt}
    pe = {k: v[perm] for k, v in pe.items()}
    sel = {"m1": rng.uniform(10, 50, 32), "q": rng.uniform(0.3, 1, 32), "z": rng.uniform(0.05, 1, 32),
           "pdraw": rng.uniform(0.5, 2, 32), "ndraw": np.full(32, 320)}
    args = inputs.model_arguments(pe, sel)
    assert args[0].shape == (nobs, nsamp) and args[4].shape == (32,) and args[8] == 320.0
    # first row belongs to the alphabetically first event, and the conversion is m1 (1+z)
    first = pe["evt"] == "GW1"
    assert np.allclose(np.sort(args[0][0]), np.sort(pe["m1"][first] * (1 + pe["z"][first])))
    with pytest.raises(ValueError):
        inputs.group_events(np.array([1, 1, 2]), np.arange(3.0))


def test_dvdzdt_table_shape_and_positivity():
    t = inputs.FlatLCDM().dVdzdt_interp()
    assert t.shape == (1024,) and t[0] == 0.0 and np.all(t[1:] > 0)

"""Host side of the fit driver (bumpcosmology_b200/run_cosmo_fit.py, the reference's run_cosmo_fit.py): table readers
and the construction of the nine model arguments.  CPU only; the sampling step is covered by tests/test_gpu_nuts.py."""
import numpy as np
import pytest

from bumpcosmology_b200 import inputs, run_cosmo_fit
from mock_tables import make_tables


def test_reference_run_configuration():
    assert (run_cosmo_fit.NMCMC, run_cosmo_fit.NCHAIN, run_cosmo_fit.RANDOM_SEED) == (1000, 4, 1652819403)


@pytest.mark.parametrize("ext", ("npz", "parquet", "csv"))
def test_tables_round_trip_through_the_readers(tmp_path, ext):
    pe, sel = make_tables(nobs=4, nsamp=16, nsel=100)
    paths = {}
    for name, tab in (("pe", pe), ("sel", sel)):
        p = tmp_path / f"{name}.{ext}"
        if ext == "npz":
            np.savez(p, **tab)
        else:
            import pandas as pd
            df = pd.DataFrame(tab)
            df.to_parquet(p) if ext == "parquet" else df.to_csv(p, index=False)
        paths[name] = str(p)
    pe2, sel2 = run_cosmo_fit.read_table(paths["pe"]), run_cosmo_fit.read_table(paths["sel"])
    a, b = inputs.model_arguments(pe, sel), inputs.model_arguments(pe2, sel2)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-14 if ext != "csv" else 1e-12)
    assert a[0].shape == (4, 16) and a[4].shape == (100,) and a[8] == 1000.0


def test_model_arguments_follow_the_reference_conversion():
    """run_cosmo_fit.py:23-30: m1d = m1 (1+z), dl = d_L(z), pdraw_cosmo = wt * dm1sqz_dm1ddqdl(m1, q, z)."""
    pe, sel = make_tables(nobs=3, nsamp=8, nsel=50)
    m1s, qs, dls, pdraws, s_m1d, s_q, s_dl, s_pd, ndraw = inputs.model_arguments(pe, sel)
    cosmo = inputs.FlatLCDM()
    z = pe["z"].reshape(3, 8)
    assert np.allclose(m1s, pe["m1"].reshape(3, 8) * (1 + z))
    assert np.allclose(dls, cosmo.luminosity_distance(z))
    jac = 1.0 / (1 + z) / (cosmo.comoving_distance(z) + (1 + z) * cosmo.dH / cosmo.efunc(z))   # weighting.py:180
    assert np.allclose(pdraws, jac)
    assert np.allclose(s_pd, sel["pdraw"] * inputs.dm1sqz_dm1ddqdl(sel["m1"], sel["q"], sel["z"], cosmo))


def test_unknown_table_format_is_rejected(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"")
    with pytest.raises(ValueError, match="unknown table format"):
        run_cosmo_fit.read_table(str(p))

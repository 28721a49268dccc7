"""`astropy.cosmology` stand-in (TEST INFRASTRUCTURE).  `pop_cosmo_model` never calls astropy; `pop_model`
(intensity_models.py:323-325) needs one theta-independent table from `Planck18.differential_comoving_volume`.
The stand-in is flat LCDM with Planck18's H0 = 67.66, Om0 = 0.30966 and NO radiation / neutrino terms, so it differs
from real astropy at the 1e-4 level — irrelevant for parity, because the table is an *input* of the path: the
oracle and the CUDA library are handed the very same numbers."""
import numpy as np

_H0, _OM = 67.66, 0.30966
_DH_GPC = 299792.458 / _H0 / 1e3


class _Quantity:
    def __init__(self, value):
        self._v = np.asarray(value, dtype=np.float64)

    def to(self, unit):
        return self

    @property
    def value(self):
        return self._v


class _Planck18:
    H0 = _H0
    Om0 = _OM

    def efunc(self, z):
        z = np.asarray(z, dtype=np.float64)
        return np.sqrt(_OM * (1 + z) ** 3 + (1 - _OM))

    def comoving_distance(self, z):
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        out = np.empty_like(z)
        for i, zi in enumerate(z.ravel()):
            x = np.linspace(0.0, zi, 4097)
            y = 1.0 / self.efunc(x)
            out.ravel()[i] = _DH_GPC * float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x)))
        return _Quantity(out)

    @property
    def hubble_distance(self):
        return _Quantity(_DH_GPC)

    def differential_comoving_volume(self, z):
        z = np.asarray(z, dtype=np.float64)
        # dense cumulative trapezoid on a log grid, then interpolate: accurate to ~1e-9, plenty for an input table
        g = np.expm1(np.linspace(0.0, np.log1p(max(float(z.max()), 1.0)), 200001))
        inv = 1.0 / self.efunc(g)
        dc = _DH_GPC * np.concatenate(([0.0], np.cumsum(0.5 * np.diff(g) * (inv[1:] + inv[:-1]))))
        dcz = np.interp(z, g, dc)
        return _Quantity(_DH_GPC * dcz ** 2 / self.efunc(z))     # Gpc^3 / sr


Planck18 = _Planck18()

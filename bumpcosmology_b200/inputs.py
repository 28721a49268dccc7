"""Input side of the hot path: what the reference's driver does between its HDF5 tables
and the model call (/root/reference/src/scripts/run_cosmo_fit.py:22-43, weighting.py:173-180).

The reference converts source-frame samples to the detector frame with astropy's Planck18, which is not available
here; `FlatLCDM` is a small numpy stand-in with the same interface for the three quantities the driver needs
(luminosity distance, comoving distance, E(z)), Planck18's H0 and Om0 by default and no radiation / neutrino terms.
Host-side, runs once per catalog; nothing here is on the per-evaluation path.
"""
import numpy as np

C_KM_S = 299792.458


class FlatLCDM:
    def __init__(self, H0=67.66, Om0=0.30966, zmax=20.0, n=65537):
        self.H0, self.Om0 = float(H0), float(Om0)
        self.dH = C_KM_S / self.H0 / 1e3                      # Gpc
        self._z = np.expm1(np.linspace(0.0, np.log1p(zmax), n))
        inv = 1.0 / self.efunc(self._z)
        self._dc = self.dH * np.concatenate(([0.0], np.cumsum(0.5 * np.diff(self._z) * (inv[1:] + inv[:-1]))))

    def efunc(self, z):
        z = np.asarray(z, dtype=np.float64)
        return np.sqrt(self.Om0 * (1 + z) ** 3 + (1 - self.Om0))

    def comoving_distance(self, z):
        return np.interp(np.asarray(z, dtype=np.float64), self._z, self._dc)

    def luminosity_distance(self, z):
        z = np.asarray(z, dtype=np.float64)
        return (1 + z) * self.comoving_distance(z)

    def differential_comoving_volume(self, z):
        """dVc/dz/dOmega in Gpc^3/sr (what pop_model tabulates, intensity_models.py:325)."""
        z = np.asarray(z, dtype=np.float64)
        return self.dH * self.comoving_distance(z) ** 2 / self.efunc(z)

    def dVdzdt_interp(self):
        """The 1024-entry table of pop_model (intensity_models.py:323-325) for `intensity_models.pop_model`."""
        zinterp = np.expm1(np.linspace(np.log1p(0), np.log1p(100), 1024))
        big = FlatLCDM(self.H0, self.Om0, zmax=100.0, n=262145) if self._z[-1] < 100 else self
        return 4 * np.pi * big.differential_comoving_volume(zinterp) / (1 + zinterp)


def dm1sqz_dm1ddqdl(m1, q, z, cosmo):
    """Jacobian d(m1_source, q, z) / d(m1_det, q, d_L)  (weighting.py:173-180)."""
    z = np.asarray(z, dtype=np.float64)
    return 1.0 / (1 + z) / (cosmo.comoving_distance(z) + (1 + z) * cosmo.dH / cosmo.efunc(z))


def detector_frame(m1, q, z, weight, cosmo=None):
    """Source-frame samples -> the model's detector-frame inputs (run_cosmo_fit.py:23-30):
    m1_det = m1 (1+z), d_L(z) [Gpc], pdraw_cosmo = weight * Jacobian.  Returns (m1_det, q, d_L, pdraw_cosmo)."""
    cosmo = cosmo or FlatLCDM()
    m1, q, z, weight = (np.asarray(x, dtype=np.float64) for x in (m1, q, z, weight))
    return m1 * (1 + z), q, cosmo.luminosity_distance(z), weight * dm1sqz_dm1ddqdl(m1, q, z, cosmo)


def group_events(evt, *columns):
    """Stack per-sample columns into [nobs, nsamp] arrays by event label, in sorted label order, as the driver's
    `groupby('evt')` loop does (run_cosmo_fit.py:32-43).  Every event must have the same number of samples."""
    evt = np.asarray(evt)
    labels, inverse, counts = np.unique(evt, return_inverse=True, return_counts=True)
    if np.any(counts != counts[0]):
        raise ValueError("events have different numbers of samples; the model needs a rectangular [nobs, nsamp] array")
    order = np.argsort(inverse, kind="stable")
    return labels, tuple(np.asarray(c, dtype=np.float64)[order].reshape(len(labels), counts[0]) for c in columns)


def model_arguments(pe, sel, cosmo=None):
    """The 9 positional arguments of `pop_cosmo_model` from the reference's two tables, given as mappings of
    columns (pandas DataFrames work): pe = {m1, q, z, wt, evt} (draw_pe_samples.py:24), sel = {m1, q, z, pdraw, ndraw}
    (draw_selection_samples.py:15)."""
    cosmo = cosmo or FlatLCDM()
    m1d, q, dl, pd = detector_frame(pe["m1"], pe["q"], pe["z"], pe["wt"], cosmo)
    _, (m1s, qs, dls, pdraws) = group_events(pe["evt"], m1d, q, dl, pd)
    s_m1d, s_q, s_dl, s_pd = detector_frame(sel["m1"], sel["q"], sel["z"], sel["pdraw"], cosmo)
    ndraw = float(np.asarray(sel["ndraw"]).ravel()[0])
    return m1s, qs, dls, pdraws, s_m1d, s_q, s_dl, s_pd, ndraw

"""Mint the fixed-cosmology golden vector (the reference's `pop_model`, intensity_models.py:313-355) from the
UNMODIFIED reference source.  Build container only:   python tests/golden/make_golden_fixed.py

Source-frame inputs are derived from the `small` synthetic catalog with the fiducial cosmology of catalogs.py; the
dVdzdt table is the one the reference model builds through the astropy stand-in (oracle/refshim/astropy)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bumpcosmology_b200 import catalogs  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def source_frame(cat):
    cosmo = catalogs._FiducialCosmology()
    z = np.interp(cat.dls, cosmo.dl, cosmo.z)
    zs = np.interp(cat.dls_sel, cosmo.dl, cosmo.z)
    return (cat.m1s_det / (1 + z), cat.qs, z, cat.pdraw, cat.m1s_det_sel / (1 + zs), cat.qs_sel, zs, cat.pdraw_sel,
            cat.Ndraw)


def main():
    cat = catalogs.make_catalog("small")
    data = source_frame(cat)
    thetas = np.vstack([catalogs.THETA_DEFAULT, catalogs.draw_prior_thetas(3, seed=21)])
    rec = {"thetas": thetas, "site_names": np.array(rr.FIXED_SITES), "Ndraw": np.float64(cat.Ndraw)}
    for k, v in zip(("m1s", "qs", "zs", "pdraw", "m1s_sel", "qs_sel", "zs_sel", "pdraw_sel"), data[:8]):
        rec[k] = v
    keys = ("loglike", "selfactor", "log_mu_sel", "neff_sel", "neff", "R", "dloglike_dsite", "dlog_mu_sel_dsite")
    outs = {k: [] for k in keys}
    for th in thetas:
        s = dict(a=th[3], b=th[4], c=th[5], mpisn=th[6], dmbhmax=th[7] - th[6], sigma=th[8], beta=th[10],
                 log_fpl=math.log(th[9]), lam=th[11], dkappa=th[12] - th[11], zp=th[13])
        r = rr.run_pop_model(s, data, R_unit=-0.5)
        for k in keys:
            outs[k].append(r[k])
        rec["dvdzdt_interp"] = r["dvdzdt_interp"]
    for k in keys:
        rec["ref_" + k] = np.array(outs[k])
    path = os.path.join(HERE, "pop_fixed_small.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path), "bytes; loglike", rec["ref_loglike"], "neff_sel", rec["ref_neff_sel"])


if __name__ == "__main__":
    main()

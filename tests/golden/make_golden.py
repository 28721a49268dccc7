"""Mint the golden vectors in this directory from the UNMODIFIED reference source.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

For each case the reference's own `pop_cosmo_model` (intensity_models.py:357-406) and its
`FlatwCDMCosmology` / `LogDNDMPISN` classes are executed through `oracle/run_reference.py` (torch-backed
stand-ins for the absent jax/numpyro, see oracle/refshim/README.md) and the outputs — factors,
deterministics, tables and the autograd gradients with respect to the 14 numpyro sample sites — are frozen
together with the inputs.  The restated oracle (`oracle/bump_oracle.py`) and the CUDA path are both tested
against these files; nothing at test time needs the reference tree.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sites_from_theta(th):
    return dict(h=th[0], Om=th[1], w=th[2], a=th[3], b=th[4], c=th[5], mpisn=th[6], dmbhmax=th[7] - th[6],
                sigma=th[8], beta=th[10], log_fpl=math.log(th[9]), lam=th[11], dkappa=th[12] - th[11], zp=th[13])


def main():
    thetas = np.vstack([THETA_DEFAULT, draw_prior_thetas(5, seed=11)])
    for cname in ("tiny", "small"):
        cat = make_catalog(cname)
        rec = {"thetas": thetas, "site_names": np.array(rr.SAMPLE_SITES)}
        for k, v in zip(("m1s_det", "qs", "dls", "pdraw", "m1s_det_sel", "qs_sel", "dls_sel", "pdraw_sel"),
                        cat.as_args()[:8]):
            rec[k] = v
        rec["Ndraw"] = np.float64(cat.Ndraw)
        keys = ("loglike", "selfactor", "log_mu_sel", "neff_sel", "neff", "R", "dloglike_dsite",
                "dlog_mu_sel_dsite", "mdNdmdVdt_fixed_qz", "dNdqdVdt_fixed_mz", "dNdVdt_fixed_mq", "hz")
        outs = {k: [] for k in keys}
        tabs = {}
        for th in thetas:
            s = sites_from_theta(th)
            r = rr.run_pop_cosmo_model(s, cat.as_args(), R_unit=0.25)
            for k in keys:
                outs[k].append(r[k])
            for k, v in rr.reference_tables(s).items():
                tabs.setdefault(k, []).append(v)
        for k in keys:
            rec["ref_" + k] = np.array(outs[k])
        for k, v in tabs.items():
            rec["tab_" + k] = np.array(v)
        path = os.path.join(HERE, f"pop_cosmo_{cname}.npz")
        np.savez_compressed(path, **rec)
        print(path, os.path.getsize(path), "bytes; loglike", rec["ref_loglike"], "neff_sel", rec["ref_neff_sel"])


if __name__ == "__main__":
    main()

"""Mint the potential-energy golden: what numpyro's NUTS integrates for the reference model.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_potential.py

`oracle/run_reference.run_potential` executes the UNMODIFIED reference `pop_cosmo_model` (its own mass_parameters /
redshift_parameters / cosmo_parameters decide which distribution every sample site gets, intensity_models.py:281-311,
398) over the torch stand-ins; the stand-in `numpyro.distributions` carries numpyro's log-densities and `biject_to`
transforms, so U(u) = -[sum log_prob(x(u)) + sum log|dx/du| + loglike + selfactor] and dU/du (torch.autograd, 15
unconstrained coordinates) come out exactly as numpyro assembles them.  Frozen here for the data of
pop_cosmo_small.npz at 6 seeded points; tests/test_potential_golden.py checks priors.py and the C++ sampler's prior
code on the CPU, and PopCosmoModel.potential / the library's potential on the GPU, against it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import run_reference as rr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = np.load(os.path.join(HERE, "pop_cosmo_small.npz"))
    data = (g["m1s_det"], g["qs"], g["dls"], g["pdraw"], g["m1s_det_sel"], g["qs_sel"], g["dls_sel"], g["pdraw_sel"],
            float(g["Ndraw"]))
    rng = np.random.default_rng(20261018)
    us = np.vstack([np.zeros(15), rng.uniform(-1.5, 1.5, (5, 15))])
    keys = ("U", "grad", "prior_U", "prior_grad", "x", "loglike", "selfactor", "R")
    out = {k: [] for k in keys}
    dists = None
    for u in us:
        r = rr.run_potential(u, data)
        for k in keys:
            out[k].append(r[k])
        dists = r["distributions"]
    rec = {"u": us, "site_names": np.array(rr.ALL_SITES), "catalog": np.array("pop_cosmo_small.npz"),
           "distributions": np.array([f"{k}: {v}" for k, v in dists.items()])}
    for k in keys:
        rec["ref_" + k] = np.array(out[k])
    path = os.path.join(HERE, "potential_small.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path), "bytes")
    print("U", rec["ref_U"])
    print("\n".join(rec["distributions"]))


if __name__ == "__main__":
    main()

"""The reference's fixed-cosmology fit driver (/root/reference/src/scripts/run_fit.py) on the CUDA hot path.

    python -m bumpcosmology_b200.run_fit --pe pe-samples.h5 --sel selection-samples.h5 --out trace.npz

Same steps as the reference script: read the two tables, stack the per-event SOURCE-frame rows (m1, q, z, wt;
run_fit.py:19-32), run NUTS(pop_model, dense_mass=True) with 1000 + 1000 steps on 4 chains and seed 3281922803
(:11-14,34-38), save the trace (:40-41).  `pop_model` (intensity_models.py:313-355) needs dVdzdt on its 1024-knot z
grid, which the reference takes from astropy's Planck18; here `inputs.FlatLCDM().dVdzdt_interp()` supplies it (same
H0, Om0, no radiation).  Chains run in parallel host threads with the Python driver (the library's C++ driver binds
pop_cosmo_model only).
"""
import argparse
import json
import time

import numpy as np

from . import inputs
from .run_cosmo_fit import NCHAIN, NMCMC, read_table

RANDOM_SEED = 3281922803   # run_fit.py:14


def fit(pe, sel, num_warmup=NMCMC, num_samples=NMCMC, num_chains=NCHAIN, seed=RANDOM_SEED, device=0, cosmo=None):
    from . import intensity_models as im, nuts
    cosmo = cosmo or inputs.FlatLCDM()
    _, (m1s, qs, zs, pdraws) = inputs.group_events(pe["evt"], pe["m1"], pe["q"], pe["z"], pe["wt"])   # :22-32
    args = (m1s, qs, zs, pdraws, np.asarray(sel["m1"], float), np.asarray(sel["q"], float),
            np.asarray(sel["z"], float), np.asarray(sel["pdraw"], float), float(np.asarray(sel["ndraw"]).ravel()[0]))
    table = cosmo.dVdzdt_interp()
    models = [im.pop_model(*args, dVdzdt_interp=table, device=device) for _ in range(num_chains)]
    try:
        t0 = time.perf_counter()
        r = nuts.run_mcmc(models, num_warmup, num_samples, num_chains, seed=seed)
        wall = time.perf_counter() - t0
    finally:
        for m in models:
            m.close()
    trace = {"site_names": np.array(im.FIXED_SITES), "posterior": r["x"], "ess_bulk": r["ess_bulk"], "rhat": r["rhat"],
             "wall_s": wall, "warmup_s": r["warmup_s"], "sampling_s": r["sampling_s"],
             "n_leapfrog": r["n_leapfrog_total"], "nobs": m1s.shape[0], "nsamp": m1s.shape[1], "nsel": len(args[4])}
    for k in r["chains"][0]["stats"]:
        trace["stat_" + k] = np.stack([np.asarray(c["stats"][k]) for c in r["chains"]])
    for k in r["chains"][0]["deterministic"]:
        trace["det_" + k] = np.stack([np.asarray(c["deterministic"][k]) for c in r["chains"]])
    return trace


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--pe", required=True, help="posterior samples: columns m1, q, z, wt, evt")
    ap.add_argument("--sel", required=True, help="found injections: columns m1, q, z, pdraw, ndraw")
    ap.add_argument("--out", default="trace.npz")
    ap.add_argument("--nmcmc", type=int, default=NMCMC)
    ap.add_argument("--nchain", type=int, default=NCHAIN)
    ap.add_argument("--seed", type=int, default=RANDOM_SEED)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    trace = fit(read_table(a.pe), read_table(a.sel), a.nmcmc, a.nmcmc, a.nchain, a.seed, a.device)
    np.savez_compressed(a.out, **trace)
    print(json.dumps({"out": a.out, "nobs": int(trace["nobs"]), "nsamp": int(trace["nsamp"]), "nsel": int(trace["nsel"]),
                      "chains": a.nchain, "draws": a.nmcmc, "wall_s": round(float(trace["wall_s"]), 3),
                      "ess_min": float(trace["ess_bulk"][:11].min()), "rhat_max": float(trace["rhat"][:11].max()),
                      "divergences": int(trace["stat_diverging"].sum())}))
    return trace


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""NUTS on a synthetic catalog with the CUDA hyperlikelihood (BASELINE.json config 2: GWTC-3 shape, 4 chains x
1000 draws after 1000 warm-up, dense mass, the reference's seed).  Prints one JSON line with ESS/s.
   python tools/run_nuts.py [--workload gwtc3] [--warmup 1000] [--samples 1000] [--chains 4]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bumpcosmology_b200 import intensity_models as im, nuts, priors
from bumpcosmology_b200.catalogs import make_catalog

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="gwtc3")
ap.add_argument("--warmup", type=int, default=1000)
ap.add_argument("--samples", type=int, default=1000)
ap.add_argument("--chains", type=int, default=4)
ap.add_argument("--seed", type=int, default=1652819403)   # run_cosmo_fit.py:19
ap.add_argument("--progress", type=int, default=0)
ap.add_argument("--out", default="")
ap.add_argument("--sequential", action="store_true", help="one context, chains one after another")
ap.add_argument("--native", action="store_true", help="the library's C++ driver (bump_nuts_chain) instead of the Python one")
a = ap.parse_args()
cat = make_catalog(a.workload)
if a.sequential:
    model = im.pop_cosmo_model(*cat.as_args())
    models = [model]
else:   # one context per chain, chains in threads (the reference's chains run in parallel as well)
    models = [im.pop_cosmo_model(*cat.as_args()) for _ in range(a.chains)]
    model = models
t0 = time.perf_counter()
r = nuts.run_mcmc(model, a.warmup, a.samples, a.chains, seed=a.seed, progress=a.progress or None, native=a.native)
wall = time.perf_counter() - t0
ess = r["ess_bulk"][:14]
x = r["x"]
line = {
    "metric": "NUTS bulk-ESS/s (min over the 14 likelihood sites)", "workload": a.workload,
    "chains": a.chains, "warmup": a.warmup, "samples": a.samples, "wall_s": wall, "sampling_s": r["sampling_s"],
    "warmup_s": r["warmup_s"], "ess_min": float(ess.min()), "ess_min_site": priors.SITE_NAMES[int(ess.argmin())],
    "ess_per_s_total": float(ess.min() / wall), "ess_per_s_sampling": float(ess.min() / r["sampling_s"]),
    "rhat_max": float(r["rhat"][:14].max()), "n_leapfrog": int(r["n_leapfrog_total"]),
    "evals_per_s": r["n_leapfrog_total"] / wall, "model_evals": sum(m.n_evals for m in models),
    "chains_in_parallel": not a.sequential, "driver": "c++ (bump_nuts_chain)" if a.native else "python (nuts.py)",
    "divergences": int(sum(c["stats"]["diverging"].sum() for c in r["chains"])),
    "mean_accept": float(np.mean([c["stats"]["accept"].mean() for c in r["chains"]])),
    "mean_depth": float(np.mean([c["stats"]["depth"].mean() for c in r["chains"]])),
    "posterior_mean": dict(zip(priors.SITE_NAMES, np.round(x.mean((0, 1)), 4).tolist())),
    "posterior_sd": dict(zip(priors.SITE_NAMES, np.round(x.std((0, 1)), 4).tolist())),
    "ess_bulk": dict(zip(priors.SITE_NAMES, np.round(r["ess_bulk"], 1).tolist())),
    "neff_sel_min": float(min(c["deterministic"]["neff_sel"].min() for c in r["chains"])),
}
print(json.dumps(line), flush=True)
if a.out:
    np.savez_compressed(a.out, x=x, site_names=np.array(priors.SITE_NAMES), ess=r["ess_bulk"], rhat=r["rhat"])

"""Kernel-time A/B on a quarter-size O5-shaped catalog (tuning only; not a bench number).
   BUMP_LIB_PATH=build/libbump_tXXX.so python tools/tune.py [--sort 0|1] [--nobs N]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bumpcosmology_b200.catalogs import THETA_DEFAULT, make_catalog
from bumpcosmology_b200.likelihood import Hyperlikelihood
ap = argparse.ArgumentParser()
ap.add_argument("--sort", type=int, default=1)
ap.add_argument("--nobs", type=int, default=1250)
ap.add_argument("--nsamp", type=int, default=10000)
ap.add_argument("--nsel", type=int, default=2_500_000)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
cat = make_catalog("o5", nobs=a.nobs, nsamp=a.nsamp, nsel=a.nsel)
t0 = time.time()
like = Hyperlikelihood(*cat.as_args(), sort=bool(a.sort))
up = time.time() - t0
like.time_evals(THETA_DEFAULT, 3)
tot, ker = like.time_evals(THETA_DEFAULT, a.iters, kernel=True)
r = like(THETA_DEFAULT)
n = cat.n_elements
print(f"lib={os.environ.get('BUMP_LIB_PATH','default')} sort={a.sort} n={n} upload_s={up:.2f} plan={like.plan()} "
      f"eval_ms={tot/a.iters:.4f} kernel_ms={ker/a.iters:.4f} ns_per_sample={ker/a.iters*1e6/n:.4f} logl={r.logl:.10f}")

"""Target for compute-sanitizer (tools/gpu_sanitize.sh): a few evaluations of the `small` catalog in every kernel mode
(default, w0-wa, fixed cosmology), launched directly (no CUDA graph, so that every kernel is instrumented), checked
against the oracle so that an instrumented run that computes garbage is noticed."""
import sys
import numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import THETA_DEFAULT, draw_prior_thetas, make_catalog
from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog, merge_partials, unpack_header
from oracle import bump_oracle as bo

cat = make_catalog("small", seed=3)
like = Hyperlikelihood(*cat.as_args(), graph=False)
for th in np.vstack([THETA_DEFAULT, draw_prior_thetas(2, seed=4)]):
    r = like(th)
    o = bo.evaluate(th, cat.as_args(), grad=True)
    assert abs(r.loglike - o["loglike"]) <= 1e-10 * max(1, abs(o["loglike"]))
    assert np.all(np.abs(r.dlog_mu_sel - o["dlog_mu_sel"]) <= 1e-10 * np.maximum(1, np.abs(o["dlog_mu_sel"])))
like.close()
wa = Hyperlikelihood(*cat.as_args(), graph=False, wa=True)
r = wa(np.concatenate([THETA_DEFAULT, [0.3]]))
assert np.isfinite(r.logl)
wa.close()
parts = []
for rank in range(2):   # two emulated ranks through the partial entry + host merge
    sh = Hyperlikelihood(*shard_catalog(cat.as_args(), rank, 2), graph=False)
    parts.append(sh.partial(THETA_DEFAULT)[0])
    sh.close()
m = unpack_header(merge_partials(np.array(parts)), 14)
assert np.isfinite(m["loglike"])
print("sanitize target ok", r.logl, m["loglike"])

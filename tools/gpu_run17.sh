#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python - <<'PY' 2>&1 | tee $out/r17_timeline.txt
import os, sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog
o5 = make_catalog("o5")
for name, cat in (("gwtc3", make_catalog("gwtc3").as_args()), ("o5/8 shard", shard_catalog(o5.as_args(), 3, 8)), ("o5", o5.as_args())):
    like = Hyperlikelihood(*cat)
    like.time_evals(THETA_DEFAULT, 20)
    n = 100
    tot, ker = like.time_evals(THETA_DEFAULT, n, kernel=True)
    tls = [like.timeline(THETA_DEFAULT) for _ in range(9)]
    med = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}
    print(name, "us/eval", round(1e3 * tot / n, 2), "stream kernel us", round(1e3 * ker / n, 2), like.plan(), flush=True)
    for k, v in med.items():
        print("   %-22s %8.2f %8.2f" % (k, v[0], v[1]))
    like.close()
PY

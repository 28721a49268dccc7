#!/bin/bash
# The host call (bump_eval through Hyperlikelihood.raw) with and without the one-graph host path, and 4 concurrent
# NUTS chains on top of each (run under gpurun):  tools/gpu_e2e_ab.sh [tag]
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-e2e}
for mode in calls onegraph calls onegraph; do
  if [ $mode = calls ]; then unset BUMP_HOST_GRAPH; else export BUMP_HOST_GRAPH=1; fi
  timeout 600 python - <<PY 2>&1 | tee -a $out/${tag}_ab.txt
import sys, time, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
row = ["$mode"]
for name in ("gwtc3", "o4", "o5"):
    like = Hyperlikelihood(*make_catalog(name).as_args())
    for _ in range(30): like.raw(THETA_DEFAULT)
    n = 2000 if name != "o5" else 200
    t0 = time.perf_counter()
    for _ in range(n): like.raw(THETA_DEFAULT)
    dt = time.perf_counter() - t0
    row.append("%s %.2f us/call" % (name, 1e6 * dt / n))
    like.close()
print(" | ".join(row), flush=True)
PY
  timeout 600 python tools/run_nuts.py --workload gwtc3_nuts --native 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('   nuts $mode', {k: round(d[k],3) for k in ('wall_s','sampling_s','ess_min','ess_per_s_total','evals_per_s','rhat_max')}, d['divergences'])" | tee -a $out/${tag}_ab.txt
done

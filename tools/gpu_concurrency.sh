#!/bin/bash
# How well do the evaluations of several contexts overlap on one GPU?  k host threads, one context (clone) each, every
# thread replaying its evaluation graph back to back with no host synchronisation in between (bump_time_evals):
# aggregate evaluations per second against k, for several shares per warp of the streaming kernel (BUMP_GPW).
set -u
out=gpurun_out; mkdir -p $out
for gpw in ${GPWS:-default 10 17}; do
  if [ $gpw = default ]; then unset BUMP_GPW; else export BUMP_GPW=$gpw; fi
  timeout 600 python - <<PY 2>&1 | tee -a $out/concurrency.txt
import sys, time, threading, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
first = Hyperlikelihood(*make_catalog("${CATALOG:-gwtc3_nuts}").as_args())
likes = [first] + [first.clone() for _ in range(7)]
for l in likes: l.time_evals(THETA_DEFAULT, 50)
row = ["gpw $gpw grid %d" % first.plan()["grid"]]
for k in (1, 2, 4, 8):
    n = 3000
    def work(i): likes[i].time_evals(THETA_DEFAULT, n)
    th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    row.append("k=%d: %.1f k evals/s (%.1f us per evaluation of one context)" % (k, k * n / dt / 1e3, 1e6 * dt / n))
print(" | ".join(row), flush=True)
for l in likes: l.close()
PY
done

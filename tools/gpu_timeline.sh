#!/bin/bash
# Per-phase timeline of one evaluation (graph mode) at the GWTC-3, O5/8-shard and O5 sizes, and the distribution of
# the streaming kernel's per-warp finish times (run under gpurun):  tools/gpu_timeline.sh [tag]
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-timeline}
timeout 900 python - <<'PY' 2>&1 | tee $out/${tag}.txt
import sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200 import _lib
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog
o5 = make_catalog("o5")
for name, cat in (("gwtc3", make_catalog("gwtc3").as_args()), ("o4", make_catalog("o4").as_args()),
                  ("o5/8 shard", shard_catalog(o5.as_args(), 3, 8)), ("o5", o5.as_args())):
    like = Hyperlikelihood(*cat)
    like.time_evals(THETA_DEFAULT, 20)
    n = 200 if name != "o5" else 40
    tot, ker = like.time_evals(THETA_DEFAULT, n, kernel=True)
    tls = [like.timeline(THETA_DEFAULT) for _ in range(9)]
    med = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}
    plan = like.plan()
    print(name, "us/eval (graph replay, back to back)", round(1e3 * tot / n, 2), "stream kernel us", round(1e3 * ker / n, 2), plan, flush=True)
    for k, v in med.items():
        print("   %-22s %8.2f %8.2f" % (k, v[0], v[1]))
    nw = min(plan["grid"] * (plan["threads"] // 32), 4096)
    buf = np.empty(nw)
    _lib.check(like.lib.bump_debug_warp_times(like._ctx, _lib.as_dp(buf), nw))
    t = buf[buf > 0]
    q = np.percentile(t[: max(1, int(0.97 * len(t)))], [0, 10, 50, 90, 100])
    print("   warp finish times, us from kernel start (0/10/50/90/100 %% of the warps with a full share):", np.round(q, 1))
    like.close()
PY

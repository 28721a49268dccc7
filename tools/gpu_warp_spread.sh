#!/bin/bash
# Where does the spread of the streaming kernel's per-warp finish times come from?  Per block (= per SM) and within a
# block, and whether the slow blocks are the same ones from one evaluation to the next (run under gpurun).
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python - <<'PY' 2>&1 | tee $out/warp_spread.txt
import sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200 import _lib
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog
o5 = make_catalog("o5")
for name, cat in (("o5/8 shard", shard_catalog(o5.as_args(), 3, 8)), ("o5", o5.as_args())):
    like = Hyperlikelihood(*cat)
    like.time_evals(THETA_DEFAULT, 20)
    plan = like.plan()
    wpb = plan["threads"] // 32
    nw = min(plan["grid"] * wpb, 4096)
    runs = []
    for rep in range(4):
        like.timeline(THETA_DEFAULT)
        buf = np.empty(nw)
        _lib.check(like.lib.bump_debug_warp_times(like._ctx, _lib.as_dp(buf), nw))
        runs.append(buf.copy())
    runs = np.array(runs)
    nb = nw // wpb
    full = nb - 4                     # the last blocks may hold short shares
    t = runs[:, : full * wpb].reshape(len(runs), full, wpb)
    print(name, plan)
    for r in range(len(runs)):
        bm = t[r].mean(axis=1)
        print("  run %d: all warps min/med/max %.1f %.1f %.1f | block means min/med/max %.1f %.1f %.1f | within-block spread (max-min) med %.1f max %.1f"
              % (r, t[r].min(), np.median(t[r]), t[r].max(), bm.min(), np.median(bm), bm.max(),
                 np.median(t[r].max(axis=1) - t[r].min(axis=1)), (t[r].max(axis=1) - t[r].min(axis=1)).max()))
    bm = t.mean(axis=2)
    print("  correlation of block means between runs:", np.round(np.corrcoef(bm)[0], 3))
    print("  warp slot (0..%d) mean finish relative to its block mean:" % (wpb - 1), np.round((t - t.mean(axis=2, keepdims=True)).mean(axis=(0, 1)), 2))
    order = np.argsort(bm.mean(axis=0))
    print("  slowest blocks:", order[-8:], np.round(bm.mean(axis=0)[order[-8:]], 1), " fastest:", order[:8], np.round(bm.mean(axis=0)[order[:8]], 1))
    np.save("gpurun_out/warp_times_%s.npy" % name.replace("/", "_").replace(" ", "_"), runs)
    like.close()
PY

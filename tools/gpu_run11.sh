#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for v in pfld; do
  lib=build/libbump_$v.so
  BUMP_LIB_PATH=$PWD/$lib timeout 300 python tools/tune.py 2>&1 | tail -1 | tee -a $out/r11_tune.txt
  BUMP_LIB_PATH=$PWD/$lib timeout 600 python - <<'PY' 2>&1 | tee -a $out/r11_timeline.txt
import os, sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
for name in ("o5",):
    cat = make_catalog(name)
    like = Hyperlikelihood(*cat.as_args())
    like.time_evals(THETA_DEFAULT, 20)
    n = 20
    tot, _ = like.time_evals(THETA_DEFAULT, n)
    tls = [like.timeline(THETA_DEFAULT) for _ in range(9)]
    med = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}
    print(os.path.basename(os.environ["BUMP_LIB_PATH"]), name, "us/eval", round(1e3 * tot / n, 2), flush=True)
    for k, v in med.items():
        print("   %-22s %8.2f %8.2f" % (k, v[0], v[1]))
    like.close()
PY
done

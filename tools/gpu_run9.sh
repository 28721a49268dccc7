#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-r9}
timeout 600 python - <<'PY' 2>&1 | tee $out/${tag}_timeline.txt
import os, sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
for name in ("gwtc3", "o4", "o5"):
    cat = make_catalog(name)
    like = Hyperlikelihood(*cat.as_args())
    like.time_evals(THETA_DEFAULT, 20)
    n = 300 if name != "o5" else 20
    tot, _ = like.time_evals(THETA_DEFAULT, n)
    tls = [like.timeline(THETA_DEFAULT) for _ in range(9)]
    med = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}
    print(name, "us/eval", round(1e3 * tot / n, 2), flush=True)
    for k, v in med.items():
        print("   %-22s %8.2f %8.2f" % (k, v[0], v[1]))
    like.close()
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_entry.py -q 2>&1 | tail -3 | tee $out/${tag}_pytest.txt

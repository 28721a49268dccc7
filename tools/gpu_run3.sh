#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for v in default cbank; do
  if [ $v = default ]; then lib=bumpcosmology_b200/libbump_b200.so; else lib=build/libbump_$v.so; fi
  echo "== $v tune"
  BUMP_LIB_PATH=$PWD/$lib timeout 300 python tools/tune.py 2>&1 | tail -1 | tee -a $out/r3_tune.txt
  echo "== $v timelines"
  BUMP_LIB_PATH=$PWD/$lib timeout 600 python - <<'PY' 2>&1 | tee -a $out/r3_timeline.txt
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
    print(os.path.basename(os.environ["BUMP_LIB_PATH"]), name, "us/eval", round(1e3 * tot / n, 2), "timeline", med, flush=True)
    like.close()
PY
done
echo "== pytest"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/r3_pytest.txt
echo "== ncu gwtc3 launch list"
cat > /tmp/g3.py <<'PY'
import sys
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
cat = make_catalog("gwtc3")
like = Hyperlikelihood(*cat.as_args(), graph=False)
for _ in range(12):
    like(THETA_DEFAULT)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r3_gwtc3_launches.csv python /tmp/g3.py > $out/r3_ncu_list.log 2>&1; echo "list rc=$?"
grep -E "prologue|stream_kernel|epilogue" $out/r3_gwtc3_launches.csv | tail -9 | awk -F'","' '{print substr($5,1,30), $NF}'

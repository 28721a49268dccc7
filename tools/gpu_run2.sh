#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
echo "== tune"
for v in default t256 t320 mm e2048; do
  if [ $v = default ]; then lib=bumpcosmology_b200/libbump_b200.so; else lib=build/libbump_$v.so; fi
  BUMP_LIB_PATH=$PWD/$lib timeout 300 python tools/tune.py 2>&1 | tail -1 | tee -a $out/r2_tune.txt
done
echo "== timelines (graph mode)"
timeout 600 python - <<'PY' 2>&1 | tee $out/r2_timeline.txt
import sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood
for name in ("gwtc3", "o4", "o5"):
    cat = make_catalog(name)
    like = Hyperlikelihood(*cat.as_args())
    like.time_evals(THETA_DEFAULT, 20)
    tot, _ = like.time_evals(THETA_DEFAULT, 200 if name != "o5" else 20)
    n = 200 if name != "o5" else 20
    tls = [like.timeline(THETA_DEFAULT) for _ in range(9)]
    med = {k: [round(float(np.median([t[k][i] for t in tls])), 2) for i in (0, 1)] for k in tls[0]}
    print(name, "us/eval (graph replay, back to back)", round(1e3 * tot / n, 2), "timeline", med)
    like.close()
PY
echo "== pytest parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee $out/r2_pytest.txt
echo "== ncu"
cmd="python tools/tune.py --iters 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/r2_launches.csv $cmd > $out/r2_ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o $out/r2_prof $cmd > $out/r2_ncu_full.log 2>&1; echo "full rc=$?"
ls -la $out/r2_prof.ncu-rep

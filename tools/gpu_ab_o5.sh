#!/bin/bash
# A/B of library builds on the O5 catalog and its 1/8 shard only (two passes):  tools/gpu_ab_o5.sh tag variant...
set -u
out=gpurun_out; mkdir -p $out
tag=$1; shift
for rep in 1 2; do
for v in "$@"; do
  if [ $v = default ]; then lib=bumpcosmology_b200/libbump_b200.so; else lib=build/libbump_$v.so; fi
  BUMP_LIB_PATH=$PWD/$lib timeout 600 python - <<'PY' 2>&1 | tee -a $out/${tag}_ab.txt
import os, sys, numpy as np
sys.path.insert(0, ".")
from bumpcosmology_b200.catalogs import make_catalog, THETA_DEFAULT
from bumpcosmology_b200.likelihood import Hyperlikelihood, shard_catalog
o5 = make_catalog("o5")
row = [os.path.basename(os.environ["BUMP_LIB_PATH"])]
for name, cat in (("o5/8", shard_catalog(o5.as_args(), 3, 8)), ("o5", o5.as_args())):
    like = Hyperlikelihood(*cat)
    like.time_evals(THETA_DEFAULT, 30)
    n = 200 if name != "o5" else 40
    tot, ker = like.time_evals(THETA_DEFAULT, n, kernel=True)
    row.append("%s %.2f us/eval (stream %.2f)" % (name, 1e3 * tot / n, 1e3 * ker / n))
    like.close()
print(" | ".join(row), flush=True)
PY
done
done

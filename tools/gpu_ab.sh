#!/bin/bash
# A/B of library builds on ONE box (box-to-box variation is a few percent): tools/gpu_ab.sh tag variant1 variant2 ...
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
for name, cat in (("gwtc3", make_catalog("gwtc3").as_args()), ("o5/8", shard_catalog(o5.as_args(), 3, 8)), ("o5", o5.as_args())):
    like = Hyperlikelihood(*cat)
    like.time_evals(THETA_DEFAULT, 30)
    n = 200 if name != "o5" else 40
    tot, ker = like.time_evals(THETA_DEFAULT, n, kernel=True)
    import hashlib
    digest = hashlib.md5(np.ascontiguousarray(like.raw(THETA_DEFAULT)).tobytes()).hexdigest()[:8]   # bitwise identity of builds
    row.append("%s %.2f us/eval (stream %.2f) %s" % (name, 1e3 * tot / n, 1e3 * ker / n, digest))
    like.close()
print(" | ".join(row), flush=True)
PY
done
done

#!/bin/bash
# 4 concurrent NUTS chains on one GPU (gwtc3_nuts catalog): wall time against the streaming kernel's share per warp
# (BUMP_GPW: groups per warp; default = spread over all SMs).  Fewer, longer blocks per chain leave SMs to the other
# chains' kernels.  Run under gpurun.
set -u
out=gpurun_out; mkdir -p $out
for gpw in default 8 10 13 17 20 26; do
  if [ $gpw = default ]; then unset BUMP_GPW; else export BUMP_GPW=$gpw; fi
  timeout 600 python tools/run_nuts.py --workload gwtc3_nuts --native 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('gpw $gpw', {k: round(d[k],3) for k in ('wall_s','sampling_s','warmup_s','ess_min','ess_per_s_total','evals_per_s','rhat_max')}, d['divergences'], d['model_evals'])" | tee -a $out/nuts_gpw.txt
done

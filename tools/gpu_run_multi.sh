#!/bin/bash
# multi-GPU run: tools/gpu_run_multi.sh N tag
set -u
N=${1:-2}; tag=${2:-m2}
out=gpurun_out; mkdir -p $out
nvidia-smi -L | tee $out/${tag}_gpus.txt
if [ "$N" = "2" ]; then
  echo "== pytest multirank"
  timeout 900 python -m pytest tests/test_gpu_multirank.py -q 2>&1 | tail -8 | tee $out/${tag}_pytest.txt
fi
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
print("result_check", d["result_check"])
print("timeline", d["run"]["timeline_us"])
print("configs", {k: (v.get("evals_per_s"), v.get("us_per_eval")) for k, v in (d.get("configs") or {}).items()})
print("clocks", d["clocks"])
PY
tail -3 $out/${tag}_bench.err
echo "== reference arm under torchrun (3 steps)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err; echo "ref rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_ref.json')); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:80])"

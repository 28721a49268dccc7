#!/bin/bash
# End-of-round evidence on one GPU: gpu tests, the bench line, the per-phase timeline, the ncu launch list and one full
# capture of the streaming kernel (run under gpurun):  tools/gpu_final.sh [tag]
set -u
out=gpurun_out; mkdir -p $out
tag=${1:-final}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $out/${tag}_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench.json
bash tools/gpu_timeline.sh ${tag}_timeline | tail -60
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-nuts"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -f -o $out/${tag}_prof $cmd > $out/${tag}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la $out/${tag}_prof.ncu-rep

#!/bin/bash
# e2e pipeline A/B: ramp-up chunk schedule + SM-multiple chunks (new defaults) vs the old fixed 6250 / 12500 chunks; predict tests first.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_predict.py -m gpu -q --no-header -x 2>&1 | tail -2
for m in knrm drmm; do
timeout 300 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${m}_ramp.json
python -c "import json; d=json.load(open('gpurun_out/bench_${m}_ramp.json')); print('$m ramp', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e_packed']['value']), d['clocks']['sm_mhz'])"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_drmmtks.csv python bench.py --model drmmtks --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_drmmtks.log 2>&1
bash scripts/gpu_ncu.sh drmmtks drmmtks_tc_kernel 14800

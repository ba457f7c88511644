#!/bin/bash
# How much do the hot zipf rows cost the gather?  KNRM bench with zipf (BASELINE workload) vs uniform ids (experiment, not a number of record).
mkdir -p gpurun_out
for ids in zipf uniform; do
CAPR_BENCH_IDS=$ids timeout 300 python bench.py --model knrm --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_knrm_ids_$ids.json
python -c "import json; d=json.load(open('gpurun_out/bench_knrm_ids_$ids.json')); print('knrm $ids', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e_packed']['value']), d['clocks']['sm_mhz'])"
done

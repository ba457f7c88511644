#!/bin/bash
# packed e2e leg (RunPredictor, allocation-stable buffers): chunk-size sweep on the BASELINE (zipf) workload; predict tests first.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_predict.py -m gpu -q --no-header -x 2>&1 | tail -2
for c in 6216 24864 6250; do
CAPR_BENCH_PACKED_CHUNK=$c timeout 300 python bench.py --model knrm --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_knrm_packed_$c.json
python -c "import json; d=json.load(open('gpurun_out/bench_knrm_packed_$c.json')); print('knrm packed chunk $c', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e_packed']['value']), d['clocks']['sm_mhz'])"
done

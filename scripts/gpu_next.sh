#!/bin/bash
# GPU round for the SURVEY §8(f) rank-1 models: parity tests, then bench lines for DRMMTKS / ConvKNRM.
mkdir -p gpurun_out
echo "== parity (hang guard 600 s)"
timeout 600 python -m pytest tests/test_gpu_next.py -m gpu -q --no-header -rf 2>&1 | tail -30
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
b() {
  timeout 400 python bench.py --model $1 --steps 5 --warmup 3 $2 > gpurun_out/bench_$1.log 2>&1
  tail -1 gpurun_out/bench_$1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']), d.get('cpu_baseline',{}).get('value'))" 2>/dev/null || { echo "$1 FAILED"; tail -15 gpurun_out/bench_$1.log | cut -c1-400; }
}
b drmmtks; b convknrm
python scripts/mma_bench.py 2>&1 | tail -30

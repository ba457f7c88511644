#!/bin/bash
mkdir -p gpurun_out
echo "== KNRM parity (hang guard 240 s)"
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -x -k "knrm" 2>&1 | tail -5
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG in tc engine"; exit 1; fi
b() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))"; }
echo "== bench"
b normal | tee gpurun_out/bench_tc_summary.txt
CAPR_DEBUG_FLAGS=0x100 b skip_pool
CAPR_DEBUG_FLAGS=0x300 b skip_pool_and_drain
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_tc.log

#!/bin/bash
mkdir -p gpurun_out
echo "== parity, all engines (hang guard 400 s)"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf 2>&1 | tail -25
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
b() { python bench.py --model $1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))"; }
echo "== bench"
b knrm tc; b drmm tc; b pacrr tc

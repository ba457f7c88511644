#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predict.py -m gpu -q --no-header -rf 2>&1 | tail -8
b() { python bench.py --model $1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))"; }
b knrm tc; b drmm tc; b pacrr tc
CAPR_SIM_ENGINE=ffma b pacrr ffma

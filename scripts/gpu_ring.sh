#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -m gpu -q --no-header -x -k "knrm or drmm or sharded or full_size" 2>&1 | tail -3
for m in knrm drmm drmmtks; do for r in 3 2; do
CAPR_SIM_RING=$r timeout 200 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m ring $r', round(d['value']), round(d['roofline']['frac'],3), round(d['e2e']['value']), round(d['e2e_packed']['value']), d['clocks']['sm_mhz'])"
done; done

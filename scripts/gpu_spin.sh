#!/bin/bash
# effect of the try_wait suspend hint on every persistent kernel
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -x -k "tc" 2>&1 | tail -3
for m in knrm drmm drmmtks pacrr; do
timeout 200 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', round(d['value']), round(d['roofline']['frac'],3), round(d['e2e']['value']))" 2>/dev/null || echo "$m failed"
done
CAPR_PACRR_CONV=ffma timeout 200 python bench.py --model pacrr --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pacrr-ffma', round(d['value']))"
timeout 300 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bert', round(d['value'],1))"

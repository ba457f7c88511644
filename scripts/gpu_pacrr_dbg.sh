#!/bin/bash
for f in 0xD70000 0x1D70000 0x1000000; do
CAPR_BENCH_NOCHECK=1 CAPR_PACRR_DEBUG=$f timeout 200 python bench.py --model pacrr --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value']))" 2>/dev/null || echo "$f failed"
done

#!/bin/bash
b() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))"; }
b zipf
CAPR_BENCH_IDS=uniform b uniform
CAPR_BENCH_IDS=uniform CAPR_DEBUG_FLAGS=0x300 b uniform_skip_pool_drain

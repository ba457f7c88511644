#!/bin/bash
export CAPR_BENCH_NO_L2PROBE=1
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'packed', round(d.get('e2e_packed',{}).get('value',0)/1e6,3))"; }
for m in drmm drmmtks; do for c in 12432 24864; do
timeout 150 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline --chunk $c 2>/dev/null | tail -1 | ex "$m chunk $c"
done; done

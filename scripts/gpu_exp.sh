#!/bin/bash
echo "== parity (hang guard)"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf -x 2>&1 | tail -12
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
b() {
  timeout 300 python bench.py --model $1 --steps 5 --warmup 3 --no-cpu-baseline > /tmp/b.log 2>&1
  tail -1 /tmp/b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))" 2>/dev/null || { echo "$1 $2 FAILED"; tail -15 /tmp/b.log | cut -c1-400; }
}
b knrm tc; b drmm tc; b pacrr tc
CAPR_DEBUG_FLAGS=0x300 b knrm skip_pool_drain
CAPR_DEBUG_FLAGS=0x100 b knrm skip_pool

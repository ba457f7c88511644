#!/bin/bash
# A/B on one box: producer id prefetch (new library) vs the previous commit's library (build/old/libcapr_old.so); parity first.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -m gpu -q --no-header -x 2>&1 | tail -1
for m in knrm drmm; do for lib in new old new old; do
if [ $lib = old ]; then export CAPR_B200_LIB=$PWD/build/old/libcapr_old.so; else unset CAPR_B200_LIB; fi
timeout 200 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m $lib', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done; done

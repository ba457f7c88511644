#!/bin/bash
# A/B of the DRMM pooling variants (CAPR_DRMM_POOL = private | atomic | noadd | skip; the register-packed variant measured in v13 was dropped): parity of every counting mode
# given, then throughput.
mkdir -p gpurun_out
MODES=${@:-private atomic}
for mode in $MODES; do
case $mode in noadd|skip) ;; *) echo "parity $mode:"; CAPR_DRMM_POOL=$mode timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -m gpu -q --no-header -x -k "drmm" 2>&1 | tail -1;; esac
CAPR_DRMM_POOL=$mode timeout 200 python bench.py --model drmm --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_drmm_${mode}.json
python -c "import json; d=json.load(open('gpurun_out/bench_drmm_${mode}.json')); print('drmm pool $mode', round(d['value']), round(d['roofline']['frac'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done

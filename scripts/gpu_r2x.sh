#!/bin/bash
# engine 2: MMA warp on scheduler 2 (warp 18, next to the producers) instead of scheduler 0 (warp 16, next to the drains) -- A/B
mkdir -p gpurun_out
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_m18.so timeout 300 python -m pytest tests/test_gpu_parity.py -q --no-header -x -k "knrm or drmm" > gpurun_out/pytest_x.log 2>&1; rc=$?; echo "m18 tests rc=$rc"; tail -2 gpurun_out/pytest_x.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_x.log | head -20; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_BENCH_NO_L2PROBE=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
for m in knrm drmm drmmtks knrm; do
timeout 100 $B --model $m 2>/dev/null | tail -1 | ex "w16 $m"
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_m18.so timeout 100 $B --model $m 2>/dev/null | tail -1 | ex "w18 $m"
done

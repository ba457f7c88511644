#!/bin/bash
# HISTORICAL: CAPR_SIM_ARRIVE / CAPR_SIM_PRODUCERS were A/B switches of experiments that were measured slower and then removed from the tree (profiles/README.md, round 2); the script documents how the numbers were taken.
# 8 producer warps in the engine-2 KNRM kernel: parity, then same-box A/B
mkdir -p gpurun_out
CAPR_SIM_PRODUCERS=8 timeout 120 python -m pytest tests/test_gpu_parity.py -q --no-header -x -rf -k "knrm" > gpurun_out/pytest_knrm_p8.log 2>&1; rc=$?; echo "knrm p8 tests rc=$rc"; tail -3 gpurun_out/pytest_knrm_p8.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_knrm_p8.log | head -20; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', 0),3), 'e2e', round(d['e2e']['value']/1e6,3), d['clocks']['sm_mhz'])"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
export CAPR_BENCH_NO_L2PROBE=1
for rep in 1 2; do
timeout 90 $B 2>/dev/null | tail -1 | ex "knrm p4"
CAPR_SIM_PRODUCERS=8 timeout 90 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_p8.json | ex "knrm p8"
done
CAPR_SIM_PRODUCERS=8 CAPR_SIM_RING=2 timeout 90 $B 2>/dev/null | tail -1 | ex "knrm p8 ring2"

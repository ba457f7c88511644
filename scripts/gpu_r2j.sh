#!/bin/bash
# HISTORICAL: CAPR_SIM_ARRIVE / CAPR_SIM_PRODUCERS were A/B switches of experiments that were measured slower and then removed from the tree (profiles/README.md, round 2); the script documents how the numbers were taken.
# engine 2 with the group-arrive stage hand-off: parity (short timeouts), then same-box A/B against the round-1 hand-off
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_engine3.py tests/test_gpu_parity.py -q --no-header -x -rf -k "knrm or engine3 or tf_dedup" > gpurun_out/pytest_knrm.log 2>&1; rc=$?; echo "knrm tests rc=$rc"; tail -4 gpurun_out/pytest_knrm.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_knrm.log | head -20; exit 0; fi
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_predict.py -q --no-header -rf -x -k "not bert and not cedr and not parade and not knrm_" > gpurun_out/pytest_family.log 2>&1; echo "family rc=$?"; tail -4 gpurun_out/pytest_family.log
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', 0),3), 'e2e', round(d['e2e']['value']/1e6,3), d['clocks']['sm_mhz'])"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
export CAPR_BENCH_NO_L2PROBE=1
for m in knrm drmm drmmtks pacrr; do
  timeout 90 $B --model $m 2>/dev/null | tail -1 | tee gpurun_out/bench_${m}_group.json | ex "$m group"
  CAPR_SIM_ARRIVE=noinc timeout 90 $B --model $m 2>/dev/null | tail -1 | ex "$m noinc"
done
CAPR_SIM_RING=2 timeout 90 $B 2>/dev/null | tail -1 | ex "knrm group ring2"
timeout 150 $B --model convknrm 2>/dev/null | tail -1 | ex "convknrm group"

#!/bin/bash
# warp-owned producer stages: parity of every KNRM-family model, then benches + trace
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine3.py -q --no-header -x -rf > gpurun_out/pytest_engine3.log 2>&1; echo "engine3 rc=$?"; tail -3 gpurun_out/pytest_engine3.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_predict.py -q --no-header -rf -x -k "not bert and not cedr and not parade" > gpurun_out/pytest_family.log 2>&1; echo "family rc=$?"; tail -4 gpurun_out/pytest_family.log
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', 0),3), 'e2e', round(d['e2e']['value']/1e6,3), d['clocks']['sm_mhz'], (d['roofline'].get('l2_gather') or {}).get('peak'))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3.json | ex knrm_e3
export CAPR_BENCH_NO_L2PROBE=1
CAPR_SIM3_QBUFS=1 timeout 300 $B 2>/dev/null | tail -1 | ex knrm_e3_q1
CAPR_KNRM_TF=0 timeout 300 $B 2>/dev/null | tail -1 | ex knrm_e3_notf
CAPR_SIM_ENGINE=tc2 timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e2.json | ex knrm_e2
for m in drmm drmmtks pacrr convknrm; do timeout 300 $B --model $m 2>/dev/null | tail -1 | tee gpurun_out/bench_$m.json | ex $m; done
timeout 300 python scripts/sim3_trace.py > gpurun_out/sim3_trace_full.txt 2> gpurun_out/sim3_trace.err
CAPR_SIM3_DEBUG=15 timeout 300 python scripts/sim3_trace.py > gpurun_out/sim3_trace_off.txt 2>> gpurun_out/sim3_trace.err
head -1 gpurun_out/sim3_trace_full.txt; head -1 gpurun_out/sim3_trace_off.txt

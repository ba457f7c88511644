#!/bin/bash
# term-frequency KNRM on engine 2: parity (short timeouts -- a protocol bug hangs), then benches
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_engine3.py -q --no-header -x -rf > gpurun_out/pytest_engine3.log 2>&1; rc=$?; echo "tf tests rc=$rc"; tail -4 gpurun_out/pytest_engine3.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_engine3.log | head -20; exit 0; fi
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_predict.py -q --no-header -rf -x -k "not bert and not cedr and not parade" > gpurun_out/pytest_family.log 2>&1; echo "family rc=$?"; tail -4 gpurun_out/pytest_family.log
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', 0),3), 'e2e', round(d['e2e']['value']/1e6,3), 'packed', round(d.get('e2e_packed',{}).get('value',0)/1e6,3), d['clocks']['sm_mhz'])"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
export CAPR_BENCH_NO_L2PROBE=1
timeout 90 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_tf_e2.json | ex knrm_tf_e2
CAPR_KNRM_TF=0 timeout 90 $B 2>/dev/null | tail -1 | ex knrm_identity_e2
CAPR_SIM_ENGINE=tc2 timeout 90 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e2.json | ex knrm_plain_e2
CAPR_KNRM_TF_ENGINE=3 timeout 90 $B 2>/dev/null | tail -1 | ex knrm_tf_e3
for m in drmm drmmtks convknrm; do timeout 120 $B --model $m 2>/dev/null | tail -1 | tee gpurun_out/bench_$m.json | ex $m; done

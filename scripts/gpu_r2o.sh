#!/bin/bash
# HISTORICAL: CAPR_SIM_ENGINE=tf (term-frequency documents on engine 2) was reverted after this A/B (profiles/README.md, round 2).
# engine 3 with 32 KB stages: parity, ablation, bench
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_engine3.py -q --no-header -x -rf > gpurun_out/pytest_engine3.log 2>&1; rc=$?; echo "engine3 rc=$rc"; tail -3 gpurun_out/pytest_engine3.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_engine3.log | head; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_BENCH_NO_L2PROBE=1 CAPR_SIM_ENGINE=tc3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
timeout 100 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3.json | ex "e3 q1 (5 stages)"
CAPR_SIM3_QBUFS=2 timeout 100 $B 2>/dev/null | tail -1 | ex "e3 q2 (3 stages)"
CAPR_KNRM_TF=0 timeout 100 $B 2>/dev/null | tail -1 | ex "e3 q1 identity"
CAPR_SIM3_STAGES=3 timeout 100 $B 2>/dev/null | tail -1 | ex "e3 q1 3 stages"
CAPR_SIM3_STAGES=4 timeout 100 $B 2>/dev/null | tail -1 | ex "e3 q1 4 stages"
CAPR_SIM_ENGINE=tc timeout 100 $B 2>/dev/null | tail -1 | ex "e2"
export CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_dbg.so CAPR_BENCH_NOCHECK=1
for dbg in 1 4 5 15; do CAPR_SIM3_DEBUG=$dbg timeout 100 $B --pairs 50000 2>/dev/null | tail -1 | ex "e3 dbg=$dbg (50k)"; done

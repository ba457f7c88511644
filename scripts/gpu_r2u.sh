#!/bin/bash
# engine 3 with balanced units of up to 160 distinct tokens (3 accumulators): parity, then A/B against engine 2 on the same box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine3.py tests/test_gpu_parity.py -q --no-header -x -rf -k "knrm or engine3 or tf_dedup or doclen" > gpurun_out/pytest_u.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/pytest_u.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_u.log | head -20; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_BENCH_NO_L2PROBE=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc"
CAPR_SIM_ENGINE=tc3 timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3"
CAPR_SIM_ENGINE=tc3 CAPR_SIM3_STAGES=3 timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3 3 stages"
CAPR_SIM_ENGINE=tc3 CAPR_SIM3_QBUFS=2 timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3 2 qbufs"
CAPR_SIM_ENGINE=tc3 CAPR_KNRM_TF=0 timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3 identity"
CAPR_SIM_ENGINE=tc3 CAPR_BENCH_IDS=uniform timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3 uniform ids"
timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc"
CAPR_SIM_ENGINE=tc3 timeout 100 $B 2>/dev/null | tail -1 | ex "knrm tc3"

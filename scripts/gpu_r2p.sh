#!/bin/bash
# HISTORICAL: CAPR_SIM_ENGINE=tf (term-frequency documents on engine 2) was reverted after this A/B (profiles/README.md, round 2).
# term-frequency documents on engine 2 (MMA warp untouched): parity, then same-box A/B against the library of commit f716625
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_engine3.py tests/test_gpu_parity.py -q --no-header -x -rf -k "knrm or engine3 or tf_dedup or doclen" > gpurun_out/pytest_knrm.log 2>&1; rc=$?; echo "knrm tests rc=$rc"; tail -3 gpurun_out/pytest_knrm.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_knrm.log | head -20; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_BENCH_NO_L2PROBE=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_base.so timeout 100 $B 2>/dev/null | tail -1 | ex "base lib knrm plain"
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_base.so timeout 100 $B --model drmm 2>/dev/null | tail -1 | ex "base lib drmm"
timeout 100 $B 2>/dev/null | tail -1 | ex "new knrm plain (tc)"
timeout 100 $B --model drmm 2>/dev/null | tail -1 | ex "new drmm"
CAPR_SIM_ENGINE=tf timeout 100 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_tf.json | ex "new knrm tf (engine 2)"
CAPR_SIM_ENGINE=tf CAPR_KNRM_TF=0 timeout 100 $B 2>/dev/null | tail -1 | ex "new knrm tf identity"
CAPR_SIM_ENGINE=tc3 timeout 100 $B 2>/dev/null | tail -1 | ex "new knrm tc3"

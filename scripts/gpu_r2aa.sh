#!/bin/bash
# engine-2 KNRM pooling with packed fp32 add / multiply (FADD2 / FMUL2): parity, then same-box A/B against the previous library
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine3.py tests/test_gpu_parity.py tests/test_gpu_next.py -q --no-header -x -k "knrm or engine3" > gpurun_out/pytest_aa.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -2 gpurun_out/pytest_aa.log
if [ $rc -ne 0 ]; then grep -n "Error\|error\|assert" gpurun_out/pytest_aa.log | head -20; exit 0; fi
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_BENCH_NO_L2PROBE=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
for i in 1 2 3; do
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_base.so timeout 100 $B 2>/dev/null | tail -1 | ex "base knrm"
timeout 100 $B 2>/dev/null | tail -1 | ex "new  knrm"
done
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_base.so timeout 100 $B --model convknrm 2>/dev/null | tail -1 | ex "base convknrm"
timeout 100 $B --model convknrm 2>/dev/null | tail -1 | ex "new  convknrm"

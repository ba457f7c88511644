#!/bin/bash
# engine-3 ablation through the debug library (results invalid, timing only)
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary --pairs 50000"
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_dbg.so CAPR_BENCH_NOCHECK=1 CAPR_BENCH_NO_L2PROBE=1
for dbg in 0 1 2 4 8 3 5 6 7 9 15; do CAPR_SIM3_DEBUG=$dbg timeout 200 $B 2>gpurun_out/err_$dbg.log | tail -1 | ex dbg=$dbg; done
for dbg in 0 1 2 4; do CAPR_KNRM_TF=0 CAPR_SIM3_DEBUG=$dbg timeout 200 $B 2>/dev/null | tail -1 | ex notf_dbg=$dbg; done

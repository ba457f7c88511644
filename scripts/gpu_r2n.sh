#!/bin/bash
# engine 2 (knrm_tc_kernel) ablation through the debug library: which role paces it?  (results invalid, timing only)
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary --pairs 50000 --skip-e2e"
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
export CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_dbg.so CAPR_BENCH_NOCHECK=1 CAPR_BENCH_NO_L2PROBE=1
for f in 0 0x100 0x200 0x300 0x400 0x800 0x700 0xb00 0xf00; do CAPR_DEBUG_FLAGS=$f timeout 100 $B 2>/dev/null | tail -1 | ex "e2 flags=$f"; done
CAPR_BENCH_IDS=uniform timeout 100 $B 2>/dev/null | tail -1 | ex "e2 uniform ids"
CAPR_SIM_RING=2 CAPR_DEBUG_FLAGS=0xf00 timeout 100 $B 2>/dev/null | tail -1 | ex "e2 ring2 flags=0xf00"

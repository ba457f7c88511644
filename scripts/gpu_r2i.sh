#!/bin/bash
# same-box A/B: library of commit f716625 (before term-frequency support in engine 2) vs the working tree
mkdir -p gpurun_out
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', 0),3), d['clocks']['sm_mhz'])"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
export CAPR_BENCH_NO_L2PROBE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,ecc.mode.current --format=csv
for rep in 1 2; do
for lib in libcapr_b200_base libcapr_b200; do
  export CAPR_B200_LIB=$PWD/capreolus_b200/$lib.so
  CAPR_SIM_ENGINE=tc2 timeout 90 $B 2>/dev/null | tail -1 | ex "$lib knrm_plain_e2"
  timeout 90 $B --model drmm 2>/dev/null | tail -1 | ex "$lib drmm"
done
done
export CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200.so
timeout 90 $B 2>/dev/null | tail -1 | ex "new knrm_tf_e2"
CAPR_KNRM_TF=0 timeout 90 $B 2>/dev/null | tail -1 | ex "new knrm_identity_e2"

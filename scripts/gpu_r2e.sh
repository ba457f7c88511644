#!/bin/bash
# A/B of the mbarrier try_wait suspend-time hint (10 ms = round 1 / none / 1 us) on engines 2 and 3 of KNRM, DRMM, PACRR and monoBERT
mkdir -p gpurun_out
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,4), 'M pairs/s  kernel_ms', round(d['roofline'].get('kernel_ms_per_launch', d['roofline'].get('forward_ms',0)),3), d['clocks']['sm_mhz'])"; }
export CAPR_BENCH_NO_L2PROBE=1
for lib in libcapr_b200 libcapr_b200_h0 libcapr_b200_h1k; do
  export CAPR_B200_LIB=$PWD/capreolus_b200/$lib.so
  B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
  timeout 300 $B 2>/dev/null | tail -1 | ex "$lib knrm_e3"
  CAPR_SIM_ENGINE=tc2 timeout 300 $B 2>/dev/null | tail -1 | ex "$lib knrm_e2"
  timeout 300 $B --model drmm 2>/dev/null | tail -1 | ex "$lib drmm"
  timeout 300 $B --model pacrr 2>/dev/null | tail -1 | ex "$lib pacrr"
  timeout 300 $B --model drmmtks 2>/dev/null | tail -1 | ex "$lib drmmtks"
  timeout 300 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | ex "$lib bert"
done

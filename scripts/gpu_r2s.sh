#!/bin/bash
# hot zipf rows: L1-cached gathers (cp.async.ca build) against the default (.cg); uniform ids as the no-hot-row bound
export PYTHONPATH=$PWD
mkdir -p gpurun_out
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3))"; }
echo "-- microbench, 148 CTAs"
for p in zipf uniform; do
  timeout 60 python scripts/gather_scaling.py $p 148
  CAPR_B200_DBG_LIB=$PWD/capreolus_b200/libcapr_b200_dbg_ca.so timeout 60 python scripts/gather_scaling.py $p 148 | sed 's/^/ca: /'
done
export CAPR_BENCH_NO_L2PROBE=1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e"
timeout 100 $B 2>/dev/null | tail -1 | ex "knrm cg zipf"
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_ca.so timeout 100 $B 2>/dev/null | tail -1 | ex "knrm ca zipf"
CAPR_BENCH_IDS=uniform timeout 100 $B 2>/dev/null | tail -1 | ex "knrm cg uniform"
CAPR_BENCH_IDS=uniform CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_ca.so timeout 100 $B 2>/dev/null | tail -1 | ex "knrm ca uniform"
CAPR_SIM_ENGINE=tc3 timeout 100 $B 2>/dev/null | tail -1 | ex "tc3 cg zipf"
CAPR_SIM_ENGINE=tc3 CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_ca.so timeout 100 $B 2>/dev/null | tail -1 | ex "tc3 ca zipf"
CAPR_SIM_ENGINE=tc3 CAPR_BENCH_IDS=uniform timeout 100 $B 2>/dev/null | tail -1 | ex "tc3 cg uniform"
timeout 100 $B --model drmm 2>/dev/null | tail -1 | ex "drmm cg zipf"
CAPR_B200_LIB=$PWD/capreolus_b200/libcapr_b200_ca.so timeout 100 $B --model drmm 2>/dev/null | tail -1 | ex "drmm ca zipf"

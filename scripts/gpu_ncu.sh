#!/bin/bash
# usage: gpu_ncu.sh <model> <kernel regex> [pairs]   -- one ncu --set full capture of the top kernel (never a bench number)
m=${1:-knrm}; k=${2:-knrm_tc_kernel}; p=${3:-14800}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${m}_full \
   python bench.py --model $m --pairs $p --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${m}.log 2>&1
ls -la gpurun_out/${m}_full.ncu-rep; tail -3 gpurun_out/ncu_${m}.log | cut -c1-300

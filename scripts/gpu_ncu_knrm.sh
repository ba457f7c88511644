#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knrm_tc_kernel -s 3 -c 1 -f -o gpurun_out/knrm_tc_full \
   python bench.py --pairs 14800 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full.log | cut -c1-300

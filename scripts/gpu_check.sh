#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list.  Usage (from the repo root, via gpurun):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --maxfail=30 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
for m in drmm pacrr; do timeout 600 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$m.log; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -25 gpurun_out/launches.csv

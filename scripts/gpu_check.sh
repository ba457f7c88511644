#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
if [ "$1" != "quick" ]; then
for m in drmm pacrr; do timeout 600 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$m.log; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -c knrm_kernel gpurun_out/launches.csv
echo "== ncu full capture (knrm_kernel, 14800 pairs)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knrm_kernel -s 3 -c 1 -f -o gpurun_out/knrm_full \
   python bench.py --pairs 14800 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi

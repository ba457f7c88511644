#!/bin/bash
# Round 2, call A: full GPU test suite + smoke + the default bench line (KNRM + monoBERT secondary) + reference arm + train mode
# + BERT launch list.   gpurun --timeout 2400 -- 'bash scripts/gpu_r2a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=30 -x -k "bert or train or next" > gpurun_out/pytest_gpu_first.log 2>&1
tail -5 gpurun_out/pytest_gpu_first.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=30 > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench (default line)"
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-1500
tail -3 gpurun_out/bench.err
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-400
echo "== bench bert"
timeout 900 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_bert.json | cut -c1-600
CAPR_BERT_CLS_ONLY=0 timeout 900 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_bert_fulltail.json | cut -c1-300
echo "== bench train"
timeout 600 python bench.py --mode train 2>/dev/null | tail -1 | tee gpurun_out/bench_train.json | cut -c1-700
echo "== ncu launch list (bert, 128 sequences)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_bert.csv \
   python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_bert.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bert.csv 2>/dev/null | head -12

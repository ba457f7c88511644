#!/bin/bash
mkdir -p gpurun_out
echo "== KNRM parity, tc engine first (hang guard 240 s)"
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -x -k "knrm and tc" 2>&1 | tail -25
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG in tc engine"; exit 1; fi
echo "== full parity file"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q --no-header -rf 2>&1 | tail -15
echo "== bench tc / ffma"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tc.log | cut -c1-1500
CAPR_SIM_ENGINE=ffma timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300

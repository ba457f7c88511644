#!/bin/bash
mkdir -p gpurun_out
echo "== smallest GEMM first (guards against a hang)"
timeout 120 python -m pytest tests/test_gpu_bert.py -m gpu -q --no-header -x -k "test_tcgen05_gemm and 128-64-64" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG in smallest GEMM"; exit 1; fi
echo "== all BERT tests"
timeout 900 python -m pytest tests/test_gpu_bert.py -m gpu -q --no-header -rf 2>&1 | tail -60 | tee gpurun_out/pytest_bert.log

#!/bin/bash
# last regression of the tree: full GPU suite, smoke, PACRR line with its fp32-pipe roofline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02d_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1; tail -1 gpurun_out/r02d_smoke.log
timeout 400 python bench.py --model pacrr --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02d_bench_pacrr.json
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_pacrr.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['fp32_pipe'])"

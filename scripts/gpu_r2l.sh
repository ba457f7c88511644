#!/bin/bash
# training heads, PACRR doc tiling, fast GELU: tests, then the BERT bench + launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_bert.py -q --no-header -x -rf > gpurun_out/pytest_train_bert.log 2>&1; echo "train+bert rc=$?"; tail -4 gpurun_out/pytest_train_bert.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q --no-header -x -rf -k "pacrr or doclen" > gpurun_out/pytest_pacrr.log 2>&1; echo "pacrr rc=$?"; tail -4 gpurun_out/pytest_pacrr.log
timeout 300 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_bert.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bert', d['value'], d['roofline']['frac'], d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_bert.csv python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_bert.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bert.csv 2>/dev/null | head -8
timeout 200 python bench.py --model pacrr --pairs 20000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pacrr', d['value'])"

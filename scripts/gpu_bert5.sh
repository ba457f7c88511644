#!/bin/bash
mkdir -p gpurun_out
echo "== encoder-model tests (hang guard 300 s)"
timeout 300 python -m pytest tests/test_gpu_bert.py tests/test_gpu_next.py -m gpu -q --no-header -rf -x -k "bert or cedr or parade" 2>&1 | tail -8
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
for v in v4 v2; do
CAPR_BERT_ATTENTION=$v timeout 600 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_bert_attn_$v.json
python -c "import json; d=json.load(open('gpurun_out/bench_bert_attn_$v.json')); print('$v', round(d['value'],1), 'pairs/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_bert.csv \
   python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_bert.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bert.csv 2>/dev/null | head -6

#!/bin/bash
mkdir -p gpurun_out
echo "== BERT tests (hang guard 300 s)"
timeout 300 python -m pytest tests/test_gpu_bert.py -m gpu -q --no-header -rf -x -k "not gemm" 2>&1 | tail -25
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
echo "== bench bert (tc attention / ffma attention)"
python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bert_tcattn.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tc-attn', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3), ' e2e', round(d['e2e']['value']))"
CAPR_BERT_ATTENTION=ffma python bench.py --model bert --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ffma-attn', round(d['value']), 'pairs/s  frac', round(d['roofline']['frac'],3))"

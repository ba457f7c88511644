#!/bin/bash
# S = Q.K^T with 3 / 2 / 1 bf16 products: error distribution on the 64-sequence BERT-base golden + speed
mkdir -p gpurun_out
for p in 3 2 1; do
  CAPR_BERT_QK_PRODUCTS=$p timeout 200 python -m pytest tests/test_gpu_bert.py -q --no-header -x -k "error_distribution or maxp_scores" > gpurun_out/pytest_bert_qk$p.log 2>&1; echo "qk=$p rc=$?"; tail -2 gpurun_out/pytest_bert_qk$p.log
  cp gpurun_out/bert_parity_base64.json gpurun_out/bert_parity_base64_qk$p.json; cp gpurun_out/bert_parity_base_p4.json gpurun_out/bert_parity_base_p4_qk$p.json
  python -c "import json; d=json.load(open('gpurun_out/bert_parity_base64_qk$p.json'))['stats']; print('  base64', d['max/scores']['max'], d['max/scores']['p99'], 'logits', d['logits']['max']); d=json.load(open('gpurun_out/bert_parity_base_p4_qk$p.json'))['stats']; print('  p4', d['max/scores']['max'], d['logits']['max'])"
  CAPR_BERT_QK_PRODUCTS=$p timeout 200 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  bert qk=$p', d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done

#!/bin/bash
# BERT: sequences per encoder call 128 vs 148 (= SM count: every GEMM's tile count divisible by the 74 clusters)
mkdir -p gpurun_out
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'pairs/s  ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
B="python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
CAPR_BERT_SEQS_PER_CALL=128 timeout 200 $B --pairs 1024 2>/dev/null | tail -1 | ex "128/call, 1024 pairs"
CAPR_BERT_SEQS_PER_CALL=148 timeout 200 $B --pairs 1036 2>/dev/null | tail -1 | ex "148/call, 1036 pairs"
CAPR_BERT_SEQS_PER_CALL=128 timeout 200 $B --pairs 1024 2>/dev/null | tail -1 | ex "128/call, 1024 pairs"
CAPR_BERT_SEQS_PER_CALL=148 timeout 200 $B --pairs 1036 2>/dev/null | tail -1 | ex "148/call, 1036 pairs"
CAPR_BERT_SEQS_PER_CALL=296 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "296/call, 1184 pairs"
CAPR_BERT_SEQS_PER_CALL=74 timeout 200 $B --pairs 1036 2>/dev/null | tail -1 | ex "74/call, 1036 pairs"
timeout 300 python -m pytest tests/test_gpu_bert.py tests/test_gpu_next.py -q --no-header -x 2>&1 | tail -2

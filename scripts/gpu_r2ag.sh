#!/bin/bash
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'pairs/s', d['clocks']['sm_mhz'])"; }
B="python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
CAPR_BERT_SEQS_PER_CALL=296 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "296/call 1184"
CAPR_BERT_SEQS_PER_CALL=592 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "592/call 1184"
CAPR_BERT_SEQS_PER_CALL=1184 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "1184/call 1184"
CAPR_BERT_SEQS_PER_CALL=296 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "296/call 1184"
CAPR_BERT_SEQS_PER_CALL=592 timeout 200 $B --pairs 1184 2>/dev/null | tail -1 | ex "592/call 1184"

#!/bin/bash
mkdir -p gpurun_out
echo "== BERT tests (hang guard 300 s)"
timeout 300 python -m pytest tests/test_gpu_bert.py -m gpu -q --no-header -rf -k "not gemm" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ "$rc" == "124" ]; then echo "HANG"; exit 1; fi
for v in v2 v1; do
CAPR_BERT_ATTENTION=$v timeout 600 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_bert_$v.json
python -c "import json; d=json.load(open('gpurun_out/bench_bert_$v.json')); print('$v', round(d['value'],1), 'pairs/s frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_bert.csv \
   python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_bert.log 2>&1
python - <<'PY'
import csv, collections
f="gpurun_out/launches_bert.csv"
tot = collections.Counter(); cnt = collections.Counter()
rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    name = r[4].split("(")[0][:60]; tot[name] += float(r[-1]); cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(8): print(f"  {v/1e6:9.3f} ms {100*v/s:5.1f}%  x{cnt[k]:4d}  avg {v/cnt[k]/1e3:8.1f} us  {k}")
PY

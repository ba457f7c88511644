#!/bin/bash
for f in 0 1 2 4 8 3 10 15; do
CAPR_BERT_ATTENTION=v2 CAPR_ATTN_DEBUG=$f timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_tc2 -s 24 -c 12 --csv --log-file gpurun_out/attn_dbg_$f.csv \
   python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/attn_dbg_$f.csv")) if len(r)>10 and r[0].isdigit()]
v=[float(r[-1]) for r in rows]
print("debug=$f", "attention us:", round(sum(v)/len(v)/1e3,1), "n=",len(v))
PY
done

#!/bin/bash
# Evidence for DRMMTKS / ConvKNRM: the fixed chunk test, full bench lines (with cpu_baseline), ncu launch lists; MMA micro-benchmark.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next.py -m gpu -q --no-header -rf 2>&1 | tail -5
for m in drmmtks convknrm; do
  timeout 600 python bench.py --model $m --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$m.json
  python -c "import json; d=json.load(open('gpurun_out/bench_$m.json')); print('$m', round(d['value']), 'pairs/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1), d['clocks'])"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$m.csv \
     python bench.py --model $m --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$m.log 2>&1
done
python - <<'PY'
import csv, collections
for m in ["drmmtks", "convknrm"]:
    f = f"gpurun_out/launches_{m}.csv"
    tot = collections.Counter(); cnt = collections.Counter()
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    for r in rows:
        name = r[4].split("(")[0][:60]; tot[name] += float(r[-1]); cnt[name] += 1
    s = sum(tot.values())
    print(f)
    for k, v in tot.most_common(8): print(f"  {v/1e6:9.3f} ms {100*v/s:5.1f}%  x{cnt[k]:4d}  {k}")
PY
python scripts/mma_bench.py 2>&1 | tail -40

#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch lists (+ optional full capture of the top kernel).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick|full]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
if [ "$1" != "quick" ]; then
for m in drmm pacrr; do timeout 600 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$m.log; done
timeout 900 python bench.py --model bert --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bert.log
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_knrm.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_bert.csv \
   python bench.py --model bert --pairs 128 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_bert.log 2>&1
python - <<'PY'
import csv, collections
for f in ["gpurun_out/launches_knrm.csv", "gpurun_out/launches_bert.csv"]:
    tot = collections.Counter(); cnt = collections.Counter()
    try:
        rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    except Exception as e:
        print(f, e); continue
    for r in rows:
        name = r[4].split("(")[0][:60]; tot[name] += float(r[-1]); cnt[name] += 1
    s = sum(tot.values())
    print(f)
    for k, v in tot.most_common(8): print(f"  {v/1e6:9.3f} ms {100*v/s:5.1f}%  x{cnt[k]:4d}  {k}")
PY
fi
if [ "$1" == "full" ]; then
echo "== ncu full capture (knrm_tc_kernel, 14800 pairs)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knrm_tc_kernel -s 3 -c 1 -f -o gpurun_out/knrm_tc_full \
   python bench.py --pairs 14800 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi

#!/bin/bash
# remaining models on the final tree (predict.py changed after gpu_v13.sh): bench lines without the CPU leg
mkdir -p gpurun_out
for m in pacrr convknrm; do timeout 300 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
for m in bert cedrknrm parade; do timeout 300 python bench.py --model $m --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
python - <<'PY'
import json
for m in ["pacrr","convknrm","bert","cedrknrm","parade"]:
    try:
        d=json.load(open(f"gpurun_out/bench_{m}.json"))
        print(m, round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(m, "FAILED", e)
PY

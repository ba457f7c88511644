#!/bin/bash
# final regression of the round on the committed tree: full GPU suite, smoke, default bench line (+ DRMM / DRMMTKS), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_knrm.json
for m in drmm drmmtks; do timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
python - <<'PY'
import json
for m in ["knrm","drmm","drmmtks"]:
    d=json.load(open(f"gpurun_out/bench_{m}.json"))
    print(m, round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "packed", round(d["e2e_packed"]["value"]), "cpu", round(d.get("cpu_baseline",{}).get("value",0),1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])
PY

#!/bin/bash
# Regression + evidence round: full GPU suite, smoke, bench lines of every model (saved as JSON), reference arm.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7 | tee gpurun_out/smoke.log
echo "== bench (headline, with cpu baseline)"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_knrm.json
for m in drmm pacrr drmmtks convknrm; do timeout 600 python bench.py --model $m --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
for m in bert cedrknrm parade; do timeout 900 python bench.py --model $m --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
python - <<'PY'
import json
for m in ["knrm","drmm","pacrr","drmmtks","convknrm","bert","cedrknrm","parade"]:
    try:
        d=json.load(open(f"gpurun_out/bench_{m}.json"))
        print(m, round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "packed", round(d.get("e2e_packed",{}).get("value",0)), "cpu", round(d.get("cpu_baseline",{}).get("value",0),1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(m, "FAILED", e)
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400

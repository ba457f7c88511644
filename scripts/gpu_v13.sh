#!/bin/bash
# v13 regression + evidence: DRMMTKS parity first (new pooling), full GPU suite, smoke, bench lines, reference arm, ncu of the DRMM kernel.
mkdir -p gpurun_out
echo "== drmmtks parity"
timeout 400 python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -m gpu -q --no-header -x -k "drmmtks or tks" 2>&1 | tail -2
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_knrm.json
for m in drmm pacrr drmmtks; do timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$m.json; done
timeout 900 python bench.py --model bert --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_bert.json
python - <<'PY'
import json
for m in ["knrm","drmm","pacrr","drmmtks","bert"]:
    try:
        d=json.load(open(f"gpurun_out/bench_{m}.json"))
        print(m, round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "cpu", round(d.get("cpu_baseline",{}).get("value",0),1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(m, "FAILED", e)
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_arm.json | cut -c1-300
echo "== ncu drmm"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_drmm.csv python bench.py --model drmm --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_drmm.log 2>&1
bash scripts/gpu_ncu.sh drmm drmm_tc_kernel 14800

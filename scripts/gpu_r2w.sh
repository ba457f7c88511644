#!/bin/bash
# final single-GPU regression of the committed tree: full GPU suite, smoke, default line (+ reference arm), monoBERT line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02b_bench_reference_arm.json
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02b_bench_default.err | tail -1 > gpurun_out/r02b_bench_default.json
timeout 600 python bench.py --model bert --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_bert.json
timeout 600 python bench.py --model cedrknrm --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_cedrknrm.json
timeout 600 python bench.py --model parade --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_parade.json
python - <<'PY'
import json
for m in ["default","reference_arm","bert","cedrknrm","parade"]:
    try:
        d=json.loads(open(f"gpurun_out/r02b_bench_{m}.json").read())
        r=d.get("roofline") or {}
        print(m, round(d["value"],1), "frac", r.get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
        if m=="default":
            s=d.get("secondary") or {}; print("  secondary", s.get("value"), (s.get("roofline") or {}).get("frac"), "e2e", (s.get("e2e") or {}).get("value")); print("  l2_gather", {k:v for k,v in (r.get("l2_gather") or {}).items() if k not in ("how","note")})
    except Exception as e:
        print(m, "FAILED", e)
PY
tail -3 gpurun_out/r02b_bench_default.err

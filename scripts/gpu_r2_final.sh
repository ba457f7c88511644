#!/bin/bash
# Round-2 evidence: full GPU suite, smoke, every bench line (default with `secondary`, reference arm, per model, train mode),
# ncu launch lists of the bench commands, one `ncu --set full` capture of knrm_tc_kernel and of gemm2_kernel, gather-rate sweep.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --maxfail=40 > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -5 gpurun_out/r02_smoke.log
echo "== reference arm, then default line"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02_bench_default.err | tail -1 > gpurun_out/r02_bench_default.json
for m in drmm pacrr drmmtks convknrm; do timeout 400 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_$m.json; done
for m in bert cedrknrm parade; do timeout 600 python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_$m.json; done
CAPR_SIM_ENGINE=tc3 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > gpurun_out/r02_bench_knrm_engine3.json
timeout 600 python bench.py --mode train 2>/dev/null | tail -1 > gpurun_out/r02_bench_train.json
python - <<'PY'
import json
for m in ["default","reference_arm","drmm","pacrr","drmmtks","convknrm","bert","cedrknrm","parade","knrm_engine3","train"]:
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{m}.json").read())
        r=d.get("roofline") or {}
        print(m, round(d["value"],1), d.get("unit"), "frac", r.get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
        if m=="default":
            s=d.get("secondary") or {}; print("  secondary", s.get("value"), (s.get("roofline") or {}).get("frac"), "l2_gather", (r.get("l2_gather") or {}).get("peak"), (r.get("l2_gather") or {}).get("logical_frac"))
    except Exception as e:
        print(m, "FAILED", e)
PY
echo "== launch lists (never bench numbers)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary-pairs 128 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bert_128seq.csv python bench.py --model bert --pairs 128 --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_default_bench.csv gpurun_out/r02_launches_bert_128seq.csv | tee gpurun_out/r02_launch_summary.txt
echo "== ncu --set full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knrm_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_knrm_tc_full python bench.py --pairs 14800 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --skip-e2e > gpurun_out/ncu_knrm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 40 -c 1 -f -o gpurun_out/r02_gemm2_full python bench.py --model bert --pairs 128 --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/ncu_gemm2.log 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/ncu_summary.py gpurun_out/r02_knrm_tc_full.ncu-rep gpurun_out/r02_knrm_tc_kernel_ncu_full.json 14800 | tail -3
python scripts/ncu_summary.py gpurun_out/r02_gemm2_full.ncu-rep gpurun_out/r02_gemm2_kernel_ncu_full.json | tail -3
echo "== gather-rate sweep (debug library): 16 KB stages in flight per SM"
timeout 200 python - <<'PY'
import json
import numpy as np, torch
from capreolus_b200 import _lib, synthetic
dbg = _lib.dbg_lib()
V, E = 30000, 300
pitch = dbg.capr_table_pitch_bf16(E)
hi = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
lo = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
n = 148 * 128 * 256
rows = torch.from_numpy(synthetic.zipf_ids(np.random.default_rng(7), (n,), V).astype(np.int32)).cuda()
out = {}
for st in (2, 3, 4, 6, 8, 10, 13):
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(dbg.capr_debug_gather_bench(hi.data_ptr(), lo.data_ptr(), V, pitch, rows.data_ptr(), n, st, torch.cuda.current_stream().cuda_stream), dbg)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out[st] = round(n * pitch * 2 * 2 / (min(ms[1:]) * 1e-3) / 1e9, 1)
    print("stages", st, "KB in flight per SM", st * 16, "GB/s", out[st])
json.dump({"unit": "GB/s", "what": "capr_debug_gather_bench: zipf rows of a [30000,%d] bf16 hi/lo table, N x 16 KB stages in flight per SM -> L2->SM gather rate (best of 3)" % pitch,
           "rate_by_stages": out}, open("gpurun_out/r02_gather_rate_sweep.json", "w"))
PY

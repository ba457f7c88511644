#!/bin/bash
# Round 2, call B: engine 3 (term-frequency KNRM pooled from TMEM) -- parity first (short timeouts: a protocol bug would hang), then A/B benches.
mkdir -p gpurun_out
echo "== engine-3 tests"
timeout 300 python -m pytest tests/test_gpu_engine3.py -q --no-header -x -rf > gpurun_out/pytest_engine3.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_engine3.log
if grep -q "passed" gpurun_out/pytest_engine3.log && ! grep -q "failed" gpurun_out/pytest_engine3.log; then
echo "== KNRM parity suite (all engines) + predict"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predict.py -q --no-header -rf -k "knrm or predict or shard or full_size or narrow" > gpurun_out/pytest_knrm.log 2>&1; echo "rc=$?"
tail -6 gpurun_out/pytest_knrm.log
echo "== bench KNRM A/B (5 steps each, no secondary / cpu baseline)"
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,3), 'M pairs/s  kernel_ms', round(d['roofline']['kernel_ms_per_launch'],3), 'e2e', round(d['e2e']['value']/1e6,3), 'packed', round(d.get('e2e_packed',{}).get('value',0)/1e6,3), d['clocks']['sm_mhz'], d['roofline'].get('l2_gather'))"; }
timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3.json | ex e3_q2
CAPR_SIM3_QBUFS=1 timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3_q1.json | ex e3_q1
CAPR_KNRM_TF=0 CAPR_BENCH_NO_L2PROBE=1 timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3_notf.json | ex e3_notf
CAPR_KNRM_TF=0 CAPR_SIM3_QBUFS=1 CAPR_BENCH_NO_L2PROBE=1 timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e3_notf_q1.json | ex e3_notf_q1
CAPR_SIM_ENGINE=tc2 CAPR_BENCH_NO_L2PROBE=1 timeout 300 $B 2>/dev/null | tail -1 | tee gpurun_out/bench_knrm_e2.json | ex e2
for st in 4 5 6; do CAPR_SIM3_STAGES=$st CAPR_BENCH_NO_L2PROBE=1 timeout 300 $B 2>/dev/null | tail -1 | ex e3_q2_stages$st; done
CAPR_BENCH_IDS=uniform CAPR_BENCH_NO_L2PROBE=1 timeout 300 $B 2>/dev/null | tail -1 | ex e3_uniform_ids
fi
echo "== CPU reference arm: allocator experiment"
timeout 200 python bench.py --impl reference --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ref plain', d['value'], d['cpu_baseline']['cores'])"
MALLOC_MMAP_THRESHOLD_=4294967296 MALLOC_TRIM_THRESHOLD_=8589934592 MALLOC_TOP_PAD_=1073741824 timeout 200 python bench.py --impl reference --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ref mallopt', d['value'])"
timeout 200 python - <<'PY'
import torch, time, sys, os
sys.argv=['x']
torch.cuda.init(); torch.zeros(1, device='cuda')
import bench
rr, model = bench.build_reranker('knrm'); state={k:v.detach().clone() for k,v in model.state_dict().items()}
step, what = bench.cpu_reference_step('knrm', state)
for _ in range(3): step()
t0=time.perf_counter(); n=0
for _ in range(6): n+=step()
print('ref after cuda init', n/(time.perf_counter()-t0))
PY

#!/bin/bash
# 2 GPUs: default line under torchrun (per-rank kernel times, kernel time before / after NCCL init, e2e with int16 ids, monoBERT secondary)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > gpurun_out/nvidia_smi_2gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_2gpu.err | tail -1 > gpurun_out/bench_2gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'packed', d.get('e2e_packed',{}).get('value'))
print('ranks', json.dumps(d.get('ranks'))[:1500])
s=d.get('secondary',{}); print('secondary', s.get('value'), s.get('roofline',{}).get('frac'), json.dumps(s.get('ranks'))[:600])
PY
tail -3 gpurun_out/bench_2gpu.err
# the same process count WITHOUT NCCL: two independent single-GPU runs side by side (is a slow kernel a property of the GPU / of co-running?)
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_solo_gpu$g.json 2>/dev/null & done; wait
for g in 0 1; do python -c "import json; d=json.loads(open('gpurun_out/bench_solo_gpu$g.json').read().strip().splitlines()[-1]); print('solo gpu$g', d['value'], d['roofline']['kernel_ms_per_launch'], d['clocks'])"; done

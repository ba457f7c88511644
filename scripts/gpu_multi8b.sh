#!/bin/bash
# 8 GPUs, final tree: the default line under torchrun (24 864-pair e2e chunks, 296 sequences per encoder call)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r02e_bench_8gpu.err | tail -1 > gpurun_out/r02e_bench_8gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_8gpu.json'))
print('KNRM value', d['value'], 'e2e', d['e2e']['value'], 'packed', d.get('e2e_packed',{}).get('value'))
r=d.get('ranks',{}); print('kernel_ms', r.get('kernel_ms'))
s=d.get('secondary',{}); print('secondary', s.get('value'), s.get('roofline',{}).get('frac'), 'e2e', (s.get('e2e') or {}).get('value'))
PY
tail -2 gpurun_out/r02e_bench_8gpu.err

#!/bin/bash
# 8 GPUs: (a) the default line under torchrun (per-rank kernel times, before / after NCCL init, e2e with int16 ids, monoBERT secondary)
#         (b) BASELINE.json configs[3] at FULL size: monoBERT, 1000 q x 1000 docs = 125 000 pairs per GPU, one NCCL all-gather of the scores
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > gpurun_out/nvidia_smi_8gpu.txt
export NCCL_DEBUG=WARN
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_8gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_8gpu.json'))
print('KNRM value', d['value'], 'e2e', d['e2e']['value'], 'packed', d.get('e2e_packed',{}).get('value'))
r=d.get('ranks',{}); print('kernel_ms', r.get('kernel_ms')); print('before nccl', r.get('kernel_ms_before_nccl_init')); print('e2e_ms', r.get('e2e_ms_total'))
s=d.get('secondary',{}); print('secondary', s.get('value'), s.get('roofline',{}).get('frac'), (s.get('ranks') or {}).get('kernel_ms'))
PY
tail -2 gpurun_out/bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --model bert --pairs 125000 --steps 1 --warmup 3 --warmup-pairs 1024 --skip-e2e --no-cpu-baseline 2> gpurun_out/bench_bert_8gpu_full.err | tail -1 > gpurun_out/bench_bert_8gpu_full.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bert_8gpu_full.json'))
print('BERT configs[3] full: value', d['value'], 'ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'issued', d['roofline']['issued_frac'])
print((d.get('ranks') or {}).get('kernel_ms')); print(d['clocks'])
PY
tail -2 gpurun_out/bench_bert_8gpu_full.err

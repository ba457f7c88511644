#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_knrm_n$N.log | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --model bert --pairs 512 --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bert_n$N.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300

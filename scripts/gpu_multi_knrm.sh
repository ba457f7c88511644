#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_multi_knrm.sh N'   (headline workload only; both arms under torchrun)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_knrm_n$N.json
python -c "import json; d=json.load(open('gpurun_out/bench_knrm_n$N.json')); print('knrm n=$N', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e_packed']['value']), d['clocks']['sm_mhz'], d['config']['parallelism'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200

#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_knrm_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('knrm n=$N', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e_packed']['value']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --model bert --pairs 512 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bert_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bert n=$N', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"

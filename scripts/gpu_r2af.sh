#!/bin/bash
ex() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_iteration'],2), 'ms/iter', d['losses'])"; }
timeout 200 python bench.py --mode train 2>/dev/null | tail -1 | ex "train"
timeout 200 python bench.py --mode train 2>/dev/null | tail -1 | ex "train"
timeout 300 python -m pytest tests/test_gpu_train.py -q --no-header -x 2>&1 | tail -2

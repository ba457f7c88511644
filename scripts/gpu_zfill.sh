#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_predict.py -m gpu -q --no-header -rf -k "not cedr and not parade" 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('knrm full-length', round(d['value']), round(d['roofline']['frac'],3))"
python - <<PY
import sys, torch, time
sys.path.insert(0, ".")
from capreolus_b200 import synthetic, reranker as R
import numpy as np
V, E = 30000, 300
def run(name, Q, D, batch):
    class Ext:
        embeddings = synthetic.embedding_table(V, E, seed=0)
        config = {"maxqlen": Q, "maxdoclen": D}
    rr = getattr(R, name)(provide={"extractor": Ext()})
    rr.build_model().to("cuda").eval()
    b = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    n = b["query"].shape[0]
    with torch.no_grad():
        for _ in range(2): rr.test(b)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): rr.test(b)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    return round(n / dt)
full800 = synthetic.throughput_batch(50000, 32, 800, V, seed=2)
print("KNRM |d|=800 full-length:", run("KNRM", 32, 800, full800), "pairs/s")
# ragged: real lengths uniform in [16, 512], queries 1..32 (the parity-set distribution at throughput size)
rng = np.random.default_rng(5)
rag = synthetic.throughput_batch(100000, 32, 512, V, seed=3)
dl = rng.integers(16, 513, size=100000); ql = rng.integers(1, 33, size=100000)
rag["posdoc"][np.arange(512)[None, :] >= dl[:, None]] = 0
rag["query"][np.arange(32)[None, :] >= ql[:, None]] = 0
for m in ("KNRM", "DRMM", "DRMMTKS", "PACRR"):
    print(m, "ragged (|d| ~ U[16,512], |q| ~ U[1,32]):", run(m, 32, 512, rag), "pairs/s")
PY

#!/usr/bin/env python
"""Is the row-gather rate limited per SM or by the shared L2 / fabric?  capr_debug_gather_bench2 (lock-step, 8 stages) on 18..148
CTAs and with zipf / uniform / sequential row ids.  Writes gpurun_out/r02_gather_scaling.json."""
import json
import os
import sys

import numpy as np
import torch

from capreolus_b200 import _lib, synthetic

V, E = 30000, 300
out = []
pattern = sys.argv[1]
grid = int(sys.argv[2])
os.environ["CAPR_GB_GRID"] = str(grid)
dbg = _lib.dbg_lib()
pitch = dbg.capr_table_pitch_bf16(E)
hi = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
lo = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
n = grid * 128 * 256
rng = np.random.default_rng(7)
if pattern == "zipf":
    r = synthetic.zipf_ids(rng, (n,), V)
elif pattern == "uniform":
    r = rng.integers(1, V, size=n)
else:
    r = (np.arange(n) % (V - 1)) + 1
rows = torch.from_numpy(r.astype(np.int32)).cuda()
x = torch.randn(1 << 26, device="cuda")
for _ in range(20):  # ramp the clocks
    x = x * 1.0001
ms = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(dbg.capr_debug_gather_bench2(hi.data_ptr(), lo.data_ptr(), V, pitch, rows.data_ptr(), n, 8, 0, torch.cuda.current_stream().cuda_stream), dbg)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = min(ms[1:])
gbs = n * pitch * 4 / (t * 1e-3) / 1e9
print(json.dumps({"rows": pattern, "ctas": grid, "gbs": round(gbs, 1), "gbs_per_sm": round(gbs / grid, 2)}))

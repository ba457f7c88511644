#!/usr/bin/env python
"""Gather rate against the number of producer warps (debug library, capr_debug_gather_bench3), zipf and uniform rows.
Writes gpurun_out/r02_gather_warps.json."""
import json

import numpy as np
import torch

from capreolus_b200 import _lib, synthetic

dbg = _lib.dbg_lib()
V, E = 30000, 300
pitch = dbg.capr_table_pitch_bf16(E)
hi = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
lo = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
n = 148 * 128 * 256
rng = np.random.default_rng(7)
x = torch.randn(1 << 26, device="cuda")
for _ in range(50):  # ramp the clocks
    x = x * 1.0001
out = []
for pattern in ("zipf", "uniform"):
    ids = synthetic.zipf_ids(rng, (n,), V) if pattern == "zipf" else rng.integers(1, V, size=n)
    rows = torch.from_numpy(ids.astype(np.int32)).cuda()
    for stages in (6, 12):
        for pw in (4, 8, 16):
            ms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(dbg.capr_debug_gather_bench3(hi.data_ptr(), lo.data_ptr(), V, pitch, rows.data_ptr(), n, stages, pw, torch.cuda.current_stream().cuda_stream), dbg)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            gbs = n * pitch * 4 / (min(ms[1:]) * 1e-3) / 1e9
            rec = {"rows": pattern, "stages": stages, "producer_warps": pw, "gbs": round(gbs, 1), "gbs_per_sm": round(gbs / 148, 1)}
            out.append(rec)
            print(rec)
json.dump({"what": "capr_debug_gather_bench3: lock-step stages of 16 KB, 4 / 8 / 16 producer warps per SM", "runs": out}, open("gpurun_out/r02_gather_warps.json", "w"), indent=1)

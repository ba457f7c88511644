#!/usr/bin/env python
"""Hand-off A/B of the gather producer (debug library, capr_debug_gather_bench2): lock-step stages (all 4 producer warps fill every
stage: 128 cp.async.mbarrier.arrive.noinc per stage) against warp-owned stages (one warp per stage: 32), with and without the
copies.  Prints GB/s and the per-SM stage period in ns; writes gpurun_out/r02_gather_handoff.json."""
import json

import numpy as np
import torch

from capreolus_b200 import _lib, synthetic

dbg = _lib.dbg_lib()
V, E = 30000, 300
pitch = dbg.capr_table_pitch_bf16(E)
atoms = (pitch + 63) // 64
hi = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
lo = torch.randn((V, pitch), device="cuda").to(torch.bfloat16)
n = 148 * 128 * 256
rows = torch.from_numpy(synthetic.zipf_ids(np.random.default_rng(7), (n,), V).astype(np.int32)).cuda()
stages_per_sm = (n // 128) * 2 * atoms / 148
out = []
for stages in (4, 8, 12):
    for mode, name in ((0, "lock-step"), (1, "warp-owned"), (2, "lock-step, no copies"), (3, "warp-owned, no copies")):
        ms = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(dbg.capr_debug_gather_bench2(hi.data_ptr(), lo.data_ptr(), V, pitch, rows.data_ptr(), n, stages, mode, torch.cuda.current_stream().cuda_stream), dbg)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = min(ms[1:])
        rec = {"stages": stages, "kb_in_flight": stages * 16, "mode": name, "ms": round(t, 3), "gbs": None if mode & 2 else round(n * pitch * 4 / (t * 1e-3) / 1e9, 1),
               "ns_per_stage_per_sm": round(t * 1e6 / stages_per_sm, 1)}
        out.append(rec)
        print(rec)
json.dump({"what": "capr_debug_gather_bench2, zipf rows of a [30000,%d] bf16 hi/lo table, 16 KB stages" % pitch, "runs": out}, open("gpurun_out/r02_gather_handoff.json", "w"), indent=1)

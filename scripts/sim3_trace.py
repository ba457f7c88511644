#!/usr/bin/env python
"""clock64 trace of CTA 0 of knrm_tc3_kernel (debug library): who waits for whom.  CAPR_SIM3_DEBUG=<bits> combines with it.
   python scripts/sim3_trace.py [pairs] > gpurun_out/sim3_trace.txt"""
import os, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["CAPR_B200_LIB"] = str(ROOT / "capreolus_b200" / "libcapr_b200_dbg.so")
import numpy as np
import torch

dev = torch.device("cuda:0")
trace = torch.zeros(5 * 1024, dtype=torch.int64, device=dev)
os.environ["CAPR_SIM3_TRACE"] = str(trace.data_ptr())
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 40
rr, model = bench.build_reranker("knrm")
model.to(dev)
gpu = {k: v.to(dev) for k, v in bench.host_batch("knrm", n, seed=2).items()}
with torch.no_grad():
    for _ in range(2):
        rr.test(gpu)
    trace.zero_()
    torch.cuda.synchronize()
    rr.test(gpu)
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(5, 1024)
names = {1: "P top", 2: "P ids_empty ok", 3: "P ids written", 4: "P q_empty ok", 5: "P Q issued", 6: "P unit start", 7: "P pair issued",
         10: "M pair start", 11: "M ids_full ok", 12: "M q_full ok", 13: "M acc_empty ok", 14: "M first d_full ok", 15: "M unit committed",
         20: "L pair start", 21: "L ids_full ok", 22: "L acc_full ok", 23: "L unit pooled", 24: "L red_empty ok", 30: "F red_full ok", 31: "F pair done"}
ev = []
for role in range(5):
    for x in t[role]:
        if x:
            ev.append((int(x) >> 8, role, int(x) & 255))
ev.sort()
t0 = ev[0][0]
# steady-state window: skip the first 6 pairs of the producer
starts = [c for c, r, tag in ev if tag == 1]
lo = starts[8] if len(starts) > 12 else t0
hi = starts[12] if len(starts) > 12 else ev[-1][0]
print(f"# producer pair period (cycles), pairs 4..: {np.diff(starts)[4:24].tolist()}")
for c, r, tag in ev:
    if lo <= c <= hi:
        print(f"{c - lo:8d}  role {r}  {names.get(tag, tag)}")

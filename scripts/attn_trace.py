"""Timeline of one CTA of attention_tc2_kernel (CAPR_ATTN_TRACE): cycles relative to the end of the CTA's prologue."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
os.environ["CAPR_ATTN_TRACE"] = str(trace.data_ptr())
rr, model = bench.build_reranker("bert")
model.to("cuda").eval()
b = {k: v.to("cuda") for k, v in bench.host_batch("bert", 128, seed=2).items()}
with torch.no_grad():
    for _ in range(2):
        rr.test(b)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(4, 64)
t0 = t[0, 0]
rel = lambda x: int(x - t0) if x else None
print("kernel entry", rel(t[0, 62]), "barriers+TMEM ready", rel(t[0, 61]), "all roles done", rel(t[0, 63]))
print("producer: prologue done 0; loads of tile t issued at", [rel(x) for x in t[0, 1:9]])
print("MMA: Q landed", rel(t[1, 0]))
for tile in range(8):
    print(f" tile {tile}: QK0 {rel(t[1,1+4*tile])} QK1 {rel(t[1,2+4*tile])} PV0 {rel(t[1,3+4*tile])} PV1 {rel(t[1,4+4*tile])}")
for g in range(2):
    print(f"softmax block {g} (S visible, max known, P buffer free, P published):")
    for tile in range(8):
        print("  tile", tile, [rel(x) for x in t[2 + g, 4 * tile:4 * tile + 4]])
    print("  last PV landed", rel(t[2 + g, 60]))

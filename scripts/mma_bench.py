import torch, sys
sys.path.insert(0, ".")
from capreolus_b200 import _lib
lib = _lib.dbg_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
print("M    N  n_mma n_acc grid  cycles/MMA")
for grid in (1, 148):
    for M in (128, 64):
        for N in (32, 64, 128, 256):
            for n_acc in (1, 2, 4):
                if n_acc * N > 512: continue
                n = 64
                _lib.check(lib.capr_debug_mma_bench(M, N, n, n_acc, 20, grid, out.data_ptr(), None))
                torch.cuda.synchronize()
                base = out[:grid].float().median().item()
                _lib.check(lib.capr_debug_mma_bench(M, N, 4 * n, n_acc, 20, grid, out.data_ptr(), None))
                torch.cuda.synchronize()
                big = out[:grid].float().median().item()
                print(f"{M:3d} {N:4d} {n:5d} {n_acc:5d} {grid:4d}  {(big - base) / (3 * n):8.1f}   (fixed {base - n * (big - base) / (3 * n):7.0f})")

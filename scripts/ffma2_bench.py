import sys, torch
sys.path.insert(0, ".")
from capreolus_b200 import _lib
lib = _lib.dbg_lib()
grid, iters = 148, 2000
scratch = torch.zeros(64 + grid * 256, device="cuda")
cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
for mode, name in ((0, "FFMA2, constant-bank scalar (UR broadcast)"), (1, "FFMA2, vector-register pair"), (2, "plain FFMA x2"),
                   (3, "FFMA2 (const) + FMNMX, 56 max per 128 FFMA2")):
    for _ in range(2):
        _lib.check(lib.capr_debug_ffma2_bench(mode, iters, grid, scratch.data_ptr(), cyc.data_ptr(), None))
    torch.cuda.synchronize()
    c = cyc.float().median().item()
    fma_per_thread = iters * 16 * 8 * 2
    # per SM: 256 threads; peak = 128 FMA/clk/SM
    print(f"{name:45s} {c:10.0f} cycles  -> {256 * fma_per_thread / c:6.1f} FMA/clk/SM (peak 128)")

"""CPU emulation of the CUDA kernels' arithmetic (prepared table, fp32 dot products, exact-match snap), used to
predict parity before spending GPU time and to document where the reference itself is rounding noise."""
import numpy as np
import torch


def emulated_similarity(table: torch.Tensor, q: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """What simtile.cuh computes: rows scaled by 1/(|row|+1e-9) once, fp32 dot, OOV exact match, snap of identical
    in-vocabulary tokens to exactly 1.0."""
    inv = 1.0 / (table.norm(dim=1) + 1e-9)
    prep = table * inv[:, None]
    V = table.shape[0]
    qr = torch.where((q > 0) & (q < V), q, torch.zeros_like(q))
    dr = torch.where((d > 0) & (d < V), d, torch.zeros_like(d))
    a, b = prep[qr], prep[dr]
    sim = a.bmm(b.transpose(1, 2))
    same = (q[:, :, None] == d[:, None, :])
    sim = torch.where(same & (q[:, :, None] > 0) & (sim > 0.5), torch.ones_like(sim), sim)
    sim = sim + (same & (q[:, :, None] < 0)).float()
    return sim

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


def emulated_topk_insert(values: np.ndarray, k: int) -> np.ndarray:
    """drmmtks.cu's register top-k: every slot is rebuilt from the OLD list, t'[i] = max(min(t[i-1], v), t[i]),
    t'[0] = max(t[0], v); columns past the end enter as -inf.  Returns the k largest values, descending (as a multiset:
    duplicates are kept, like torch.topk)."""
    t = np.full(k, -np.inf, dtype=np.float32)
    for v in np.asarray(values, dtype=np.float32):
        old = t.copy()
        t[0] = max(old[0], v)
        for i in range(1, k):
            t[i] = max(min(old[i - 1], v), old[i])
    return t


def emulated_drmm_counts(sim_row: np.ndarray, doc_ids: np.ndarray, qid: int, nbins: int, slices: int = 8) -> np.ndarray:
    """drmm.cu's counting for one query row: arithmetic bin guess fixed up against the exact fp32 torch.linspace bounds,
    identical in-vocabulary tokens forced into the last regular bin, pads in no bin, one byte histogram per column slice
    (pooling warp) summed at the end.  Returns int counts [nbins + 1] (last slot = 0.999 < s < 1.001, DRMM.py:66)."""
    ub = torch.linspace(-1, 1, nbins + 1)[1:].numpy().astype(np.float32)
    hist = np.zeros((slices, nbins + 1), dtype=np.uint8)
    scale = np.float32(0.5 * nbins)
    for c, (v, did) in enumerate(zip(np.asarray(sim_row, dtype=np.float32), doc_ids)):
        if did == 0:
            continue
        g0 = int(max(0, min(int(np.floor((v + np.float32(1.0)) * scale)), nbins - 1)))
        b = g0 + (1 if v >= ub[g0] else 0) - (1 if (g0 > 0 and v < ub[max(g0 - 1, 0)]) else 0)
        if v == np.float32(1.0) and qid > 0 and qid == did:
            b = nbins - 1
        w = (c // 32) % slices
        if b < nbins:
            hist[w, b] += 1
        if np.float32(0.999) < v < np.float32(1.001):
            hist[w, nbins] += 1
    return hist.astype(np.int64).sum(axis=0)

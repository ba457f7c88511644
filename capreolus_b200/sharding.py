"""Multi-GPU scoring: candidate lists shard over pairs, one all-gather of fp32 scores at the end.

Every (query, doc) pair is scored independently (no cross-pair op in ``KNRM.py:39-55``, ``DRMM.py:101-116``,
``PACRR.py:43-54``, ``ptBERTMaxP.py:67-96``), so rank r of R scores the contiguous slice
``[r*ceil(N/R), (r+1)*ceil(N/R))`` with the same kernels and the same replicated weights; the only
communication is one ``all_gather`` of ``ceil(N/R)`` floats per rank (NCCL on GPUs, gloo in the CPU tests).
The reference's PyTorch path is single-device (``trainer/pytorch.py:203,328``); this is the data-parallel
predict loop it lacks.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of ``n`` items owned by ``rank``; the last shards may be short or empty."""
    per = math.ceil(n / world) if n else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Slice every tensor / list of a trainer batch dict (``trainer/pytorch.py:342``) along the batch axis."""
    n = len(next(v for v in batch.values() if torch.is_tensor(v)))
    lo, hi = shard_bounds(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_scores(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank score vectors (each padded to ceil(N/R)) and strip the padding -> ``[n_total]`` on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = math.ceil(n_total / world)
    padded = torch.zeros(per, dtype=torch.float32, device=local.device)
    padded[: local.shape[0]] = local.float()
    out = torch.empty(per * world, dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]


class ShardedScorer:
    """``scorer.test(batch)`` == ``reranker.test(batch)`` bit for bit, with the pairs split over the ranks of ``group``."""

    def __init__(self, reranker, group=None):
        self.reranker = reranker
        self.group = group

    def test(self, batch: dict) -> torch.Tensor:
        if not dist.is_initialized():
            return self.reranker.test(batch)
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        n = len(next(v for v in batch.values() if torch.is_tensor(v)))
        local = self.reranker.test(shard_batch(batch, rank, world))
        return gather_scores(local.view(-1), n, self.group)

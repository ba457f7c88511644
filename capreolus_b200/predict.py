"""Host-side predict loop: host batches in, host scores out.

Mirrors the hot loop of ``PytorchTrainer.predict`` (``capreolus/trainer/pytorch.py:338-348``): move the batch to
the device, ``reranker.test(batch)``, bring the scores back.  The reference does this synchronously per batch
(``.cpu()`` every iteration); here host->device copies of chunk i+1 run on a copy stream while chunk i is being
scored, and the scores of all chunks come back with one synchronisation at the end.
"""
from __future__ import annotations

import torch


class PinnedBatch:
    """A host batch (dict of tensors / lists as the DataLoader collates them) moved to pinned memory once."""

    def __init__(self, batch: dict):
        self.tensors = {k: (v.contiguous().pin_memory() if not v.is_pinned() else v) for k, v in batch.items() if torch.is_tensor(v)}
        self.n = len(next(iter(self.tensors.values())))
        self.bytes_per_item = sum(v[0].numel() * v.element_size() for v in self.tensors.values()) if self.n else 0


class PipelinedPredictor:
    """Double-buffered H2D / score / D2H pipeline around ``reranker.test``."""

    def __init__(self, reranker, device, chunk: int = 16384):
        self.reranker, self.device, self.chunk = reranker, torch.device(device), int(chunk)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = None
        self._out_host = None

    def _buffers(self, pb: PinnedBatch):
        sig = tuple((k, v.dtype, tuple(v.shape[1:])) for k, v in pb.tensors.items())
        if self._bufs is None or self._bufs[0] != sig:
            bufs = [{k: torch.empty((self.chunk,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device) for k, v in pb.tensors.items()}
                    for _ in range(2)]
            self._bufs = (sig, bufs)
        if self._out_host is None or self._out_host.shape[0] < pb.n:
            self._out_host = torch.empty(pb.n, dtype=torch.float32).pin_memory()
        return self._bufs[1]

    @torch.no_grad()
    def predict(self, pb: PinnedBatch) -> torch.Tensor:
        """Scores for every item of ``pb`` as a pinned host fp32 tensor ``[n]`` (valid after this call returns)."""
        bufs = self._buffers(pb)
        main = torch.cuda.current_stream(self.device)
        n_chunks = (pb.n + self.chunk - 1) // self.chunk
        copied = [torch.cuda.Event() for _ in range(n_chunks)]
        scored = [torch.cuda.Event() for _ in range(n_chunks)]
        self.copy_stream.wait_stream(main)
        for i in range(n_chunks):
            lo, hi = i * self.chunk, min(pb.n, (i + 1) * self.chunk)
            buf = bufs[i % 2]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(scored[i - 2])  # buffer reuse
                for k, v in pb.tensors.items():
                    buf[k][: hi - lo].copy_(v[lo:hi], non_blocking=True)
                copied[i].record(self.copy_stream)
            main.wait_event(copied[i])
            scores = self.reranker.test({k: v[: hi - lo] for k, v in buf.items()})
            self._out_host[lo:hi].copy_(scores.view(-1), non_blocking=True)
            scored[i].record(main)
        main.synchronize()
        return self._out_host[: pb.n]

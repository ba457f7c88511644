"""Drop-in counterparts of the torch pieces of ``capreolus/reranker/common.py`` (the TF twins are out of scope).

Same class names and parameter names as the reference so checkpoints interchange
(``kernels.kernels.{i}.mu`` ...), but every ``forward`` enqueues a hand-written sm_100a kernel through the C
ABI (``capreolus_b200/_lib.py``) instead of running torch ops.
"""
from __future__ import annotations

import ctypes
import os

import torch

from capreolus_b200 import _lib


#: cosine-tile engine for inference: "tc" = tcgen05 tensor cores (table as bf16 hi/lo planes; engine 2, csrc/simtc.cuh), "tc3" =
#: KNRM on engine 3 (term-frequency documents, cosines pooled straight from tensor memory, csrc/simtc3.cuh + knrm_tc3.cu:
#: parity-green, measured slower than engine 2 -- DESIGN.md; the other models stay on engine 2), "ffma" = fp32 CUDA cores.
#: Shapes the tensor-core kernels do not cover (maxdoclen > 1024, emb dim > 320, ...) use "ffma".
ENGINE = os.environ.get("CAPR_SIM_ENGINE", "tc")
DEBUG_FLAGS = int(os.environ.get("CAPR_DEBUG_FLAGS", "0"), 0)  # profiling only (CAPR_DEBUG_SKIP_*): results invalid


def use_tensor_cores(D: int, E: int, max_doclen: int = 1024) -> bool:
    """KNRM / DRMM / DRMMTKS pool column-additively over 256-doc units, so their tensor-core kernels take maxdoclen <= 1024 (the
    reference extractor's default is 800); PACRR needs the whole 512-column tile with its halo (``max_doclen=512``)."""
    return ENGINE in ("tc", "tc3") and D <= max_doclen and E <= 320


def create_emb_layer(weights, non_trainable=True):
    """``create_emb_layer`` (capreolus/reranker/common.py:279-288): nn.Embedding holding the extractor's table."""
    layer = torch.nn.Embedding(*weights.shape)
    layer.load_state_dict({"weight": torch.as_tensor(weights, dtype=torch.float32)})
    layer.weight.requires_grad = not non_trainable
    return layer


class PreparedTable:
    """Device-resident, L2-normalised, pitch-padded copy of an ``nn.Embedding`` weight (``capr_table_prepare``).

    Rebuilt lazily when the weight moves to another device or is modified in place (``finetune=True`` steps).
    Not part of ``state_dict`` -- it is derived data."""

    def __init__(self, embedding: torch.nn.Embedding):
        self.embedding = embedding
        self._key = None
        self._table = None
        self._key16 = None
        self._planes = None

    def get_bf16(self):
        """(hi, lo) bf16 planes ``[V, pitch64]`` of the normalised table for the tensor-core engine (``capr_table_prepare_bf16``)."""
        w = self.embedding.weight
        _lib.require_cuda(w)
        key = (w.data_ptr(), w._version, w.device)
        if key != self._key16:
            V, E = w.shape
            pitch = _lib.lib().capr_table_pitch_bf16(E)
            src = w.detach().contiguous()
            hi = torch.empty((V, pitch), dtype=torch.bfloat16, device=w.device)
            lo = torch.empty((V, pitch), dtype=torch.bfloat16, device=w.device)
            _lib.check(_lib.lib().capr_table_prepare_bf16(src.data_ptr(), V, E, hi.data_ptr(), lo.data_ptr(), pitch, _lib.current_stream(w.device)))
            self._planes, self._key16 = (hi, lo), key
        return self._planes

    @property
    def pitch(self) -> int:
        return _lib.lib().capr_table_pitch(self.embedding.weight.shape[1])

    def get(self) -> torch.Tensor:
        w = self.embedding.weight
        _lib.require_cuda(w)
        key = (w.data_ptr(), w._version, w.device)
        if key != self._key:
            V, E = w.shape
            src = w.detach().contiguous()
            table = torch.empty((V, self.pitch), dtype=torch.float32, device=w.device)
            _lib.check(_lib.lib().capr_table_prepare(src.data_ptr(), V, E, table.data_ptr(), self.pitch, _lib.current_stream(w.device)))
            self._table, self._key = table, key
        return self._table


def _ids(t: torch.Tensor) -> torch.Tensor:
    """Token ids as contiguous int64 (the reference extractor emits np.long, embedtext.py:146-147)."""
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


class SimilarityMatrix(torch.nn.Module):
    """``SimilarityMatrix`` (capreolus/reranker/common.py:143-182): OOV exact match + cosine matrix, pads zeroed.

    The scoring kernels fuse this producer and never materialise the matrix; this module exists for API parity
    and for tests (``capr_simmat_forward``)."""

    def __init__(self, embedding):
        super().__init__()
        self.embedding = embedding
        self.padding = 0
        self._prepared = PreparedTable(embedding)

    def forward(self, query_tok, doc_tok):
        _lib.require_cuda(query_tok, doc_tok)
        q, d = _ids(query_tok), _ids(doc_tok)
        B, Q = q.shape
        assert d.shape[0] == B
        D = d.shape[1]
        table = self._prepared.get()
        out = torch.empty((B, Q, D), dtype=torch.float32, device=q.device)
        _lib.check(_lib.lib().capr_simmat_forward(q.data_ptr(), d.data_ptr(), B, Q, D, table.data_ptr(), table.shape[0], table.shape[1],
                                                 out.data_ptr(), _lib.current_stream(q.device)))
        return out


class RbfKernel(torch.nn.Module):
    """Parameter holder with the reference's names (capreolus/reranker/common.py:224-234)."""

    def __init__(self, initial_mu, initial_sigma, requires_grad=True):
        super().__init__()
        self.mu = torch.nn.Parameter(torch.tensor(initial_mu), requires_grad=requires_grad)
        self.sigma = torch.nn.Parameter(torch.tensor(initial_sigma), requires_grad=requires_grad)


class RbfKernelBank(torch.nn.Module):
    """``RbfKernelBank`` (capreolus/reranker/common.py:237-250) as a parameter holder; the Gaussian kernels are
    evaluated inside ``capr_knrm_forward``.  ``stacked()`` returns the K mus / sigmas as two contiguous device
    vectors (cached until a parameter is modified in place, e.g. by an optimizer step)."""

    def __init__(self, mus=None, sigmas=None, dim=1, requires_grad=True):
        super().__init__()
        self.dim = dim
        self.kernels = torch.nn.ModuleList([RbfKernel(m, s, requires_grad=requires_grad) for m, s in zip(mus, sigmas)])
        self._cache_key, self._cache = None, None

    def count(self):
        return len(self.kernels)

    def stacked(self, differentiable=False):
        mus, sigmas = [k.mu for k in self.kernels], [k.sigma for k in self.kernels]
        if differentiable:
            return torch.stack(mus), torch.stack(sigmas)
        key = tuple((p.data_ptr(), p._version) for p in mus + sigmas)
        if key != self._cache_key:
            with torch.no_grad():
                self._cache = (torch.stack(mus).float().contiguous(), torch.stack(sigmas).float().contiguous())
            self._cache_key = key
        return self._cache


class _PairLoss(torch.autograd.Function):
    """A pairwise loss kernel that returns the loss and d loss / d score in one launch (``capr_pair_hinge`` / ``capr_pair_softmax``)."""

    @staticmethod
    def forward(ctx, pos, neg, entry):
        _lib.require_cuda(pos, neg)
        pos, neg = pos.contiguous().float(), neg.contiguous().float()
        B = pos.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=pos.device)
        gpos, gneg = torch.empty_like(pos), torch.empty_like(neg)
        _lib.check(getattr(_lib.lib(), entry)(pos.data_ptr(), neg.data_ptr(), B, loss.data_ptr(), gpos.data_ptr(), gneg.data_ptr(),
                                              _lib.current_stream(pos.device)))
        ctx.save_for_backward(gpos, gneg)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gpos, gneg = ctx.saved_tensors
        return g * gpos, g * gneg, None


def pair_softmax_loss(pos_neg_scores, *args, **kwargs):
    """``pair_softmax_loss`` (capreolus/reranker/common.py:96-98): mean(1 - softmax(stack([pos, neg], 1), 1)[:, 0])."""
    return _PairLoss.apply(pos_neg_scores[0], pos_neg_scores[1], "capr_pair_softmax")


def pair_hinge_loss(pos_neg_scores, *args, **kwargs):
    """``pair_hinge_loss`` (capreolus/reranker/common.py:7,101-103): MarginRankingLoss(margin=1, mean), target +1."""
    return _PairLoss.apply(pos_neg_scores[0], pos_neg_scores[1], "capr_pair_hinge")


def device_pointer_array(tensors):
    """Host array of device pointers (for capr_pacrr_forward's conv_w / conv_b)."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr

"""KNRM behind the reference's module API (``capreolus/reranker/KNRM.py``), scored by ``capr_knrm_forward``.

``KNRM_class`` keeps the reference's submodule / parameter names (``kernels.kernels.{i}.mu|sigma``,
``embedding.weight``, ``combine.{0,2}.weight|bias``), so ``Reranker.save_weights`` / ``load_weights`` files
interchange with the reference's.  ``forward`` has the same signature and returns the same ``[B,1]`` tensor.
"""
from __future__ import annotations

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker import common
from capreolus_b200.reranker.common import PreparedTable, RbfKernelBank, SimilarityMatrix, _ids, create_emb_layer

SCORETANH = 1  # CAPR_KNRM_SCORETANH


class _KnrmFeatures(torch.autograd.Function):
    """log soft-TF features ``[B,K]`` with the closed-form gradient to mu / sigma (kernel K4).

    The kernel emits, per pair and kernel, dR_k/dmu_k and dR_k/dsigma_k; backward is a [B,K] contraction."""

    @staticmethod
    def forward(ctx, mu, sigma, q, d, table):
        B, Q = q.shape
        D = d.shape[1]
        K = mu.shape[0]
        feats = torch.empty((B, K), dtype=torch.float32, device=q.device)
        stats = torch.empty((B, 2, K), dtype=torch.float32, device=q.device)
        mu_c, sigma_c = mu.detach().float().contiguous(), sigma.detach().float().contiguous()
        _lib.check(_lib.lib().capr_knrm_forward(
            q.data_ptr(), d.data_ptr(), B, Q, D, table.data_ptr(), table.shape[0], table.shape[1], mu_c.data_ptr(), sigma_c.data_ptr(), K,
            None, None, 0, None, None, 0, None, feats.data_ptr(), stats.data_ptr(), _lib.current_stream(q.device)))
        ctx.save_for_backward(stats)
        return feats

    @staticmethod
    def backward(ctx, g):
        (stats,) = ctx.saved_tensors
        g = g.float()
        return (g * stats[:, 0, :]).sum(dim=0), (g * stats[:, 1, :]).sum(dim=0), None, None, None


class KNRM_class(nn.Module):
    """``KNRM_class`` (capreolus/reranker/KNRM.py:13-55)."""

    def __init__(self, extractor, config):
        super(KNRM_class, self).__init__()
        self.p = config
        mus = [-0.9, -0.7, -0.5, -0.3, -0.1, 0.1, 0.3, 0.5, 0.7, 0.9, 1.0]
        sigmas = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.001]
        self.kernels = RbfKernelBank(mus, sigmas, dim=1, requires_grad=config["gradkernels"])
        self.embedding = create_emb_layer(extractor.embeddings, non_trainable=not self.p["finetune"])
        self.simmat = SimilarityMatrix(self.embedding)
        self._prepared = self.simmat._prepared
        self._tf_ws = None  # scratch of the term-frequency pre-pass (capr_tf_workspace_bytes), grown on demand

        channels = 1
        if config["singlefc"]:
            combine_steps = [nn.Linear(self.kernels.count() * channels, 1)]
        else:
            combine_steps = [nn.Linear(self.kernels.count() * channels, 30), nn.Tanh(), nn.Linear(30, 1)]
        if config["scoretanh"]:
            combine_steps.append(nn.Tanh())
        self.combine = nn.Sequential(*combine_steps)

    def get_embedding(self, toks):
        return self.embedding(toks)

    def kernel_features(self, doctoks, querytoks):
        """The ``[B,K]`` tensor the reference feeds to ``self.combine`` (KNRM.py:53); differentiable in mu / sigma."""
        q, d = _ids(querytoks), _ids(doctoks)
        mu, sigma = self.kernels.stacked(differentiable=True)
        return _KnrmFeatures.apply(mu, sigma, q, d, self._prepared.get())

    def forward(self, doctoks, querytoks, query_idf):
        _lib.require_cuda(doctoks, querytoks)
        needs_grad = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())  # eval mode: inference kernel
        if needs_grad:
            if self.embedding.weight.requires_grad:
                # finetune=True (KNRM.py:23-24,68): the gradient reaches the embedding table, which the fused kernels treat as
                # frozen -> torch restatement of the forward on the device (train_heads.py); the prepared tables are rebuilt from
                # the updated weight at the next inference call (PreparedTable keys on the weight's version counter)
                from capreolus_b200.reranker import train_heads

                return train_heads.knrm_forward(self, doctoks, querytoks)
            return self.combine(self.kernel_features(doctoks, querytoks))
        q, d = _ids(querytoks), _ids(doctoks)
        B, Q = q.shape
        D = d.shape[1]
        mu, sigma = self.kernels.stacked()
        fc1 = self.combine[0]
        hidden = 0 if self.p["singlefc"] else fc1.out_features
        fc2 = None if self.p["singlefc"] else self.combine[2]
        scores = torch.empty((B, 1), dtype=torch.float32, device=q.device)
        E = self.embedding.weight.shape[1]
        if common.use_tensor_cores(D, E) and mu.shape[0] <= 16 and common.ENGINE == "tc3":
            # engine 3 (opt-in): term-frequency documents, pooling straight from tensor memory (csrc/knrm_tc3.cu)
            hi, lo = self._prepared.get_bf16()
            lib = _lib.lib()
            need = lib.capr_tf_workspace_bytes(B, D)
            if self._tf_ws is None or self._tf_ws.numel() < need or self._tf_ws.device != q.device:
                self._tf_ws = torch.empty(max(need, 256), dtype=torch.uint8, device=q.device)
            _lib.check(lib.capr_knrm_forward_tf(
                q.data_ptr(), d.data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], E, hi.shape[1], mu.data_ptr(), sigma.data_ptr(),
                mu.shape[0], fc1.weight.data_ptr(), fc1.bias.data_ptr(), hidden, _lib.ptr(fc2.weight if fc2 is not None else None),
                _lib.ptr(fc2.bias if fc2 is not None else None), SCORETANH if self.p["scoretanh"] else 0, scores.data_ptr(), None,
                self._tf_ws.data_ptr(), self._tf_ws.numel(), _lib.current_stream(q.device)))
            return scores
        if common.use_tensor_cores(D, E) and mu.shape[0] <= 16:
            hi, lo = self._prepared.get_bf16()
            _lib.check(_lib.lib().capr_knrm_forward_tc(
                q.data_ptr(), d.data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], E, hi.shape[1], mu.data_ptr(), sigma.data_ptr(),
                mu.shape[0], fc1.weight.data_ptr(), fc1.bias.data_ptr(), hidden, _lib.ptr(fc2.weight if fc2 is not None else None),
                _lib.ptr(fc2.bias if fc2 is not None else None), (SCORETANH if self.p["scoretanh"] else 0) | common.DEBUG_FLAGS, scores.data_ptr(), None,
                _lib.current_stream(q.device)))
            return scores
        table = self._prepared.get()
        _lib.check(_lib.lib().capr_knrm_forward(
            q.data_ptr(), d.data_ptr(), B, Q, D, table.data_ptr(), table.shape[0], table.shape[1], mu.data_ptr(), sigma.data_ptr(),
            mu.shape[0], fc1.weight.data_ptr(), fc1.bias.data_ptr(), hidden, _lib.ptr(fc2.weight if fc2 is not None else None),
            _lib.ptr(fc2.bias if fc2 is not None else None), SCORETANH if self.p["scoretanh"] else 0, scores.data_ptr(), None, None,
            _lib.current_stream(q.device)))
        return scores


@Reranker.register
class KNRM(Reranker):
    """Chenyan Xiong, Zhuyun Dai, Jamie Callan, Zhiyuan Liu, and Russell Power. 2017. End-to-End Neural Ad-hoc Ranking with Kernel Pooling. In SIGIR'17."""

    module_name = "KNRM"

    config_spec = [
        ConfigOption("gradkernels", True, "backprop through mus and sigmas"),
        ConfigOption("scoretanh", False, "use a tanh on the prediction as in paper (True) or do not use a nonlinearity (False)"),
        ConfigOption("singlefc", True, "use single fully connected layer as in paper (True) or 2 fully connected layers (False)"),
        ConfigOption("finetune", False, "fine tune the embedding layer"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = KNRM_class(self.extractor, self.config)
        return self.model

    def score(self, d):
        query_idf = d["query_idf"]
        query_sentence = d["query"]
        pos_sentence, neg_sentence = d["posdoc"], d["negdoc"]
        m = self.model
        if m.training and torch.is_grad_enabled() and not m.embedding.weight.requires_grad and pos_sentence.shape == neg_sentence.shape:
            # training step (KNRM.py:87-94 scores the positive and the negative documents with two forward calls): one launch / one autograd node
            # for both -- pairs are independent, so the scores are the same; the iteration is host-bound (bench.py --mode train)
            B = pos_sentence.shape[0]
            both = m(torch.cat([pos_sentence, neg_sentence]), torch.cat([query_sentence, query_sentence]), None).view(-1)
            return [both[:B], both[B:]]
        return [
            m(pos_sentence, query_sentence, query_idf).view(-1),
            m(neg_sentence, query_sentence, query_idf).view(-1),
        ]

    def test(self, d):
        query_idf = d["query_idf"]
        query_sentence = d["query"]
        pos_sentence = d["posdoc"]
        return self.model(pos_sentence, query_sentence, query_idf).view(-1)

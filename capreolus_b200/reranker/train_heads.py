"""Differentiable heads for TRAINING the embedding-id rerankers other than default KNRM (round 2).

``PytorchTrainer.single_train_iteration`` (``capreolus/trainer/pytorch.py:76-122``) calls ``reranker.score(batch)`` in train mode for
every reranker.  BASELINE.json keeps one training configuration in scope -- KNRM's pairwise-hinge loop, which has its own kernels
(closed-form d/dmu, d/dsigma statistics, ``capr_knrm_forward``) -- but a drop-in must not raise when the trainer hands it another
model.  For those, ``model.train()`` + grad mode selects the path below:

* everything UPSTREAM of the trainable parameters is still the CUDA engine (no gradient is needed there because the embedding table
  is frozen): DRMM's matching histogram, DRMMTKS' top-k cosines and PACRR's cosine matrix come from ``capr_drmm_forward`` /
  ``capr_drmmtks_forward_tc`` / ``capr_simmat_forward``;
* the parameterised tail is the reference's own op sequence in torch ON THE DEVICE, so autograd provides the backward: DRMM's
  ``ffw`` / term gate / output layer (``DRMM.py:83-116``), DRMMTKS' (``DRMMTKS.py:32-63``), PACRR's Conv2d / max / top-k / MLP
  (``PACRR.py:43-82``);
* ConvKNRM's Conv1d encoders sit upstream of the cosine (``ConvKNRM.py:43-77``) and KNRM with ``finetune=True`` trains the table
  itself (``KNRM.py:23-24,68``), so there the whole forward is the torch restatement (device tensors, no CPU, no oracle import).

Inference (``model.eval()``) never comes here.  These heads are library ops, not hand-written kernels: training of these models is
outside the benchmarked hot path and is provided for API completeness.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rbf_bank(sim: torch.Tensor, mus, sigmas) -> torch.Tensor:
    """``RbfKernelBank.forward`` (capreolus/reranker/common.py:232-250): stack of exp(-0.5 (s-mu)^2 / sigma / sigma) on dim 1."""
    return torch.stack([torch.exp(-0.5 * (sim - m) * (sim - m) / s / s) for m, s in zip(mus, sigmas)], dim=1)


def cosine_matrix(a_emb: torch.Tensor, b_emb: torch.Tensor, q_tok: torch.Tensor, d_tok: torch.Tensor, padding: int = 0) -> torch.Tensor:
    """cosine of every (query, doc) vector pair with <pad> rows / columns zeroed (common.py:160-167, 203-216)."""
    a_denom = a_emb.norm(p=2, dim=2)[:, :, None] + 1e-9
    b_denom = b_emb.norm(p=2, dim=2)[:, None, :] + 1e-9
    sim = a_emb.bmm(b_emb.permute(0, 2, 1)) / (a_denom * b_denom)
    sim = torch.where((q_tok == padding)[:, :, None], torch.zeros_like(sim), sim)
    return torch.where((d_tok == padding)[:, None, :], torch.zeros_like(sim), sim)


def similarity_matrix(embedding: torch.nn.Embedding, q_tok: torch.Tensor, d_tok: torch.Tensor) -> torch.Tensor:
    """``SimilarityMatrix.forward`` (common.py:170-182), differentiable in the embedding weight (KNRM finetune=True)."""
    qn, dn = q_tok.clamp(max=0), d_tok.clamp(max=0)
    exact = (qn[:, :, None] == dn[:, None, :]).float()
    exact = torch.where((qn == 0)[:, :, None], torch.zeros_like(exact), exact)
    exact = torch.where((dn == 0)[:, None, :], torch.zeros_like(exact), exact)
    qp, dp = q_tok.clamp(min=0), d_tok.clamp(min=0)
    return exact + cosine_matrix(embedding(qp), embedding(dp), qp, dp)


def kernel_pool(simmats: torch.Tensor, mus, sigmas) -> torch.Tensor:
    """KNRM.py:41-53 / ConvKNRM.py:62-75 on ``simmats [B, VIEWS, Q, D]`` -> ``[B, K*VIEWS]`` (index k*VIEWS + view)."""
    kernels = rbf_bank(simmats, mus, sigmas)  # [B, K, VIEWS, Q, D]
    B, K, VIEWS, Q, D = kernels.shape
    kernels = kernels.reshape(B, K * VIEWS, Q, D)
    sims = simmats.reshape(B, 1, VIEWS, Q, D).expand(B, K, VIEWS, Q, D).reshape(B, K * VIEWS, Q, D)
    result = kernels.sum(dim=3)
    mask = sims.sum(dim=3) != 0.0
    result = torch.where(mask, (result + 1e-6).log(), mask.float())
    return result.sum(dim=2)


def knrm_forward(model, doctoks, querytoks):
    """``KNRM_class.forward`` (KNRM.py:39-55) with a trainable embedding table."""
    sim = similarity_matrix(model.embedding, querytoks.long(), doctoks.long())
    mus = [k.mu for k in model.kernels.kernels]
    sigmas = [k.sigma for k in model.kernels.kernels]
    return model.combine(kernel_pool(sim[:, None], mus, sigmas))


def term_gate(model, query_idf, q_mask, query_tok=None):
    """``_term_gate`` (DRMM.py:83-99 / DRMMTKS.py:32-48): softmax over the query of w.idf (IDF) or w.E[q] (TV), pads at -1e7."""
    atten_mask = (1 - q_mask) * -1e7
    if model.gate_type == "IDF":
        logits = model.gates(query_idf.float()[:, :, None]).reshape(q_mask.shape) + atten_mask
    else:
        logits = model.gates(model.embedding(query_tok)).reshape(q_mask.shape) + atten_mask  # raw (un-clamped) ids, DRMM.py:109
    return F.softmax(logits, dim=1)


def drmm_forward(model, hist, query_tok, query_idf):
    """DRMM.py:101-116 downstream of ``_hist_map`` (``hist [B,Q,nbins+1]`` from capr_drmm_forward)."""
    B, Q = query_tok.shape
    z = model.ffw(hist).reshape(B, Q)
    gate = term_gate(model, query_idf, (query_tok != 0).float(), query_tok)
    return model.output_layer((gate * z).sum(dim=-1, keepdim=True))


def drmmtks_forward(model, topk, query_tok, query_idf):
    """DRMMTKS.py:50-63 downstream of ``torch.topk`` (``topk [B,Q,k]`` from capr_drmmtks_forward_tc)."""
    B, Q = query_tok.shape
    z = model.ffw(topk).reshape(B, Q)
    gate = term_gate(model, query_idf, (query_tok != 0).float(), query_tok)
    return model.output_layer((gate * z).sum(dim=-1, keepdim=True))


def pacrr_forward(model, simmat, query_idf):
    """PACRR.py:43-82 downstream of ``self.simmat`` (``simmat [B,Q,D]`` from capr_simmat_forward)."""
    B, Q, D = simmat.shape
    scores = []
    for ng in model.ngrams:
        x = simmat.reshape(B, 1, Q, D)
        if ng.shape != 1:
            x = F.pad(x, (0, ng.shape - 1, 0, ng.shape - 1))
        conv = F.relu(ng.conv(x))
        top_filters, _ = conv.max(dim=1)
        top_toks, _ = top_filters.topk(ng.k, dim=2)
        scores.append(top_toks.reshape(B, Q, ng.k))
    if model.p["idf"]:
        scores.append(F.softmax(query_idf.float(), dim=1).view(B, Q, 1))  # PACRR.py:48-50 (softmax over the query axis)
    scores = torch.cat(scores, dim=2)
    return model.combine(scores.reshape(B, -1))


def convknrm_forward(model, sentence, query_sentence):
    """``ConvKNRM_class.forward`` (ConvKNRM.py:43-77): Conv1d n-gram encoders -> stacked cosine matrices -> kernel pooling."""
    q, d = query_sentence.long(), sentence.long()
    a_emb, b_emb = model.embeddings(q), model.embeddings(d)
    a_reps, b_reps = [], []
    for pad, conv in zip(model.padding, model.convs):
        a_reps.append(conv[0](pad(a_emb.permute(0, 2, 1))).permute(0, 2, 1))
        b_reps.append(conv[0](pad(b_emb.permute(0, 2, 1))).permute(0, 2, 1))
    if model.p["crossmatch"]:
        pairs = [(a, b) for a in a_reps for b in b_reps]
    else:
        pairs = list(zip(a_reps, b_reps))
    simmats = torch.stack([cosine_matrix(a, b, q, d, model.simmat.padding) for a, b in pairs], dim=1)  # [B, VIEWS, Q, D]
    mus = [k.mu for k in model.kernels.kernels]
    sigmas = [k.sigma for k in model.kernels.kernels]
    return model.combine(kernel_pool(simmats, mus, sigmas))

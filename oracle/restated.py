"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference reranker scoring path (torch, fp32).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product (``capreolus_b200/``) never does and fails loudly
when its CUDA library is missing.

Why a restatement exists at all: the real reference modules can be executed verbatim in the build
container (``oracle/refshim.py``) but ``/root/reference`` does not travel to the GPU box.  This file
is the travelling checker.  It is *pinned*: ``tests/test_oracle.py`` checks every function below
against ``tests/golden/*.npz``, which ``oracle/make_goldens.py`` produced by running the unmodified
reference classes (commit 789288c) on the seeded inputs of ``capreolus_b200/synthetic.py``.
(The reference's own test-suite holds no numeric assertion for this path -- SURVEY.md §8c -- so
reference-generated goldens are the only pin available.)

Each function keeps the reference's *operation sequence* (it materialises the same [B,K,Q,D]
temporaries, loops over bins, etc.), because it also serves as the CPU baseline ("port") that
``bench.py`` times on the GPU box's host cores.  Parameters are passed as a flat ``dict`` that uses
the reference's ``state_dict`` key names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

KNRM_MUS = [-0.9, -0.7, -0.5, -0.3, -0.1, 0.1, 0.3, 0.5, 0.7, 0.9, 1.0]  # reranker/KNRM.py:20
KNRM_SIGMAS = [0.1] * 10 + [0.001]  # reranker/KNRM.py:21


# --------------------------------------------------------------------------------------------------
# shared: similarity matrix   (reranker/common.py:143-182)
# --------------------------------------------------------------------------------------------------
def _zero_pads(sim, q, d):
    """reranker/common.py:149-153 -- zero every cell whose query or doc id equals the padding id 0."""
    sim = torch.where((q == 0)[:, :, None], torch.zeros_like(sim), sim)
    return torch.where((d == 0)[:, None, :], torch.zeros_like(sim), sim)


def similarity_matrix(table: torch.Tensor, q: torch.Tensor, d: torch.Tensor, exact_cosines: bool = False) -> torch.Tensor:
    """reranker/common.py:170-182.  ``q [B,Q]``, ``d [B,D]`` int64 -> ``[B,Q,D]`` fp32.

    ids > 0 in-vocab, 0 pad, < 0 OOV (l.174).  Exact-match part: ids clamped to <= 0 so only identical
    negative ids match (l.155-158,179).  Cosine part: ids clamped to >= 0; the raw dot product is divided
    by the product of (norm + 1e-9) (l.160-167,180).

    ``exact_cosines=True`` evaluates the same expression with a float64 copy of the table and returns float64
    ("the reference in exact arithmetic").  The fp32 self-cosine of a token, a.a / (|a|+1e-9)^2, lands on either
    side of 1.0 by rounding; consumers that threshold at 1.0 (DRMM's last bin, ``DRMM.py:63-65``) or differentiate a
    sigma=0.001 kernel at mu=1.0 (KNRM) amplify that noise to O(1) -- see DESIGN.md "Exact matches".
    """
    if exact_cosines:
        table = table.double()
    qn, dn = q.clamp(max=0), d.clamp(max=0)
    exact = _zero_pads((qn[:, :, None] == dn[:, None, :]).to(table.dtype), qn, dn)
    qp, dp = q.clamp(min=0), d.clamp(min=0)
    a, b = F.embedding(qp, table), F.embedding(dp, table)
    a_den = a.norm(p=2, dim=2)[:, :, None] + 1e-9
    b_den = b.norm(p=2, dim=2)[:, None, :] + 1e-9
    cos = _zero_pads(a.bmm(b.permute(0, 2, 1)) / (a_den * b_den), qp, dp)
    return exact + cos


# --------------------------------------------------------------------------------------------------
# KNRM   (reranker/KNRM.py:39-55, reranker/common.py:224-250)
# --------------------------------------------------------------------------------------------------
def rbf_bank(sim: torch.Tensor, mus, sigmas) -> torch.Tensor:
    """reranker/common.py:232-234,249-250: exp(-0.5*(s-mu)^2/sigma/sigma) stacked on dim 1."""
    out = []
    for mu, sigma in zip(mus, sigmas):
        adj = sim - mu
        out.append(torch.exp(-0.5 * adj * adj / sigma / sigma))
    return torch.stack(out, dim=1)


def knrm_kernel_features(sim: torch.Tensor, mus, sigmas) -> torch.Tensor:
    """reranker/KNRM.py:41-53 -> ``[B,K]`` log soft-TF features.

    Kernel sums run over ALL doc positions (padded ones have s=0); a query row counts iff its row of
    ``sim`` does not sum to exactly 0 (l.51); masked rows contribute 0 (l.52)."""
    kern = rbf_bank(sim, mus, sigmas)  # [B,K,Q,D]
    B, K, Q, D = kern.shape
    sim_k = sim.reshape(B, 1, Q, D).expand(B, K, Q, D).reshape(B, K, Q, D)
    soft_tf = kern.sum(dim=3)
    live = sim_k.sum(dim=3) != 0.0
    feats = torch.where(live, (soft_tf + 1e-6).log(), live.float())
    return feats.sum(dim=2)


def knrm_params_from_state(state: dict) -> dict:
    K = len([k for k in state if k.startswith("kernels.kernels.") and k.endswith(".mu")])
    return {
        "mus": [state[f"kernels.kernels.{i}.mu"] for i in range(K)],
        "sigmas": [state[f"kernels.kernels.{i}.sigma"] for i in range(K)],
    }


def knrm_forward(state: dict, table, doc, query, query_idf=None, singlefc=True, scoretanh=False) -> torch.Tensor:
    """``KNRM_class.forward(doctoks, querytoks, query_idf)`` (reranker/KNRM.py:39-55) -> ``[B,1]``."""
    kp = knrm_params_from_state(state)
    feats = knrm_kernel_features(similarity_matrix(table, query, doc), kp["mus"], kp["sigmas"])
    x = F.linear(feats, state["combine.0.weight"], state["combine.0.bias"])
    if not singlefc:  # reranker/KNRM.py:31
        x = F.linear(torch.tanh(x), state["combine.2.weight"], state["combine.2.bias"])
    if scoretanh:  # reranker/KNRM.py:32-33
        x = torch.tanh(x)
    return x


# --------------------------------------------------------------------------------------------------
# DRMM   (reranker/DRMM.py:41-116)
# --------------------------------------------------------------------------------------------------
def drmm_histogram(sim: torch.Tensor, doc: torch.Tensor, nbins=29, hist_type="LCH") -> torch.Tensor:
    """reranker/DRMM.py:58-81 -> ``[B,Q,nbins+1]``.

    Padded doc columns are pushed to +1e7 (l.59); count-below per upper bound ``linspace(-1,1,nbins+1)[1:]``
    (l.63-65); last slot = #(0.999 < s < 1.001) (l.66); slots nbins-1..1 are differenced (l.68-69);
    +1 (l.71); then NH / LCH / CH."""
    d_mask = (doc != 0).to(sim.dtype)
    s = sim + (1 - d_mask[:, None, :]) * 1e7
    hist = torch.zeros(s.shape[0], s.shape[1], nbins + 1, dtype=torch.float)
    bounds = torch.linspace(-1, 1, nbins + 1)[1:]
    for i in range(nbins):
        hist[:, :, i] = (s < bounds[i]).sum(dim=-1)
    hist[:, :, -1] = ((s > 0.999) * (s < 1.001)).sum(dim=-1)
    for i in range(nbins - 1, 0, -1):
        hist[:, :, i] -= hist[:, :, i - 1]
    hist += 1
    if hist_type == "NH":
        hist = hist / hist.sum(dim=-1)[:, :, None]
    elif hist_type == "LCH":
        hist = torch.log(hist)
    elif hist_type != "CH":
        raise ValueError("histType should be 'CH', 'NH', or 'LCH'")
    return hist


def drmm_forward(state: dict, table, doc, query, query_idf, nbins=29, hist_type="LCH", gate_type="IDF",
                 exact_cosines: bool = False) -> torch.Tensor:
    """``DRMM_class.forward(sentence, query_sentence, query_idf)`` (reranker/DRMM.py:101-116) -> ``[B,1]``.

    ``exact_cosines``: bin float64 cosines (everything downstream of the integer counts stays fp32)."""
    B, Q = query.shape
    q_mask = (query != 0).float()
    hist = drmm_histogram(similarity_matrix(table, query, doc, exact_cosines), doc, nbins, hist_type)
    z = torch.tanh(F.linear(hist, state["ffw.0.weight"], state["ffw.0.bias"]))
    z = torch.tanh(F.linear(z, state["ffw.2.weight"], state["ffw.2.bias"])).reshape(B, Q)  # l.106
    neg = (1 - q_mask) * -1e7  # l.89
    if gate_type == "IDF":
        logits = F.linear(query_idf.float()[:, :, None], state["gates.weight"]).reshape(B, Q) + neg  # l.92
    elif gate_type == "TV":
        logits = F.linear(F.embedding(query, table), state["gates.weight"]).reshape(B, Q) + neg  # l.94,109
    else:
        raise ValueError("gateType should be either IDF or TV")
    gate = F.softmax(logits, dim=1)
    x = (gate * z).sum(dim=-1, keepdim=True)
    return F.linear(x, state["output_layer.weight"], state["output_layer.bias"])


# --------------------------------------------------------------------------------------------------
# PACRR   (reranker/PACRR.py:43-82)
# --------------------------------------------------------------------------------------------------
def pacrr_ngram_topk(sim: torch.Tensor, weight, bias, n: int, k: int) -> torch.Tensor:
    """``PACRRConvMax2dModule.forward`` (reranker/PACRR.py:73-82): pad bottom/right by n-1, Conv2d(1->F,n),
    ReLU, max over filters, top-k over the doc axis (descending) -> ``[B,Q,k]``."""
    B, Q, D = sim.shape
    x = sim.reshape(B, 1, Q, D)
    if n != 1:
        x = F.pad(x, (0, n - 1, 0, n - 1), value=0.0)
    conv = F.relu(F.conv2d(x, weight, bias))
    best_filter, _ = conv.max(dim=1)
    top, _ = best_filter.topk(k, dim=2)
    return top.reshape(B, Q, k)


def pacrr_forward(state: dict, table, doc, query, query_idf, mingram=1, maxgram=3, kmax=2, idf=True,
                  nonlinearity="relu") -> torch.Tensor:
    """``PACRR_class.forward(sentence, query_sentence, query_idf)`` (reranker/PACRR.py:43-54) -> ``[B,1]``.

    ``PACRR.py:49`` intends ``softmax(query_idf.reshape(B,Q,1), dim=1)`` (a softmax over the query axis)."""
    B, Q = query.shape
    sim = similarity_matrix(table, query, doc)
    feats = [
        pacrr_ngram_topk(sim, state[f"ngrams.{i}.conv.weight"], state[f"ngrams.{i}.conv.bias"], n, kmax)
        for i, n in enumerate(range(mingram, maxgram + 1))
    ]
    if idf:
        feats.append(F.softmax(query_idf.float().reshape(B, Q, 1), dim=1))
    x = torch.cat(feats, dim=2).reshape(B, -1)
    act = {"relu": F.relu, "tanh": torch.tanh, "none": lambda t: t}[nonlinearity]
    x = act(F.linear(x, state["linear1.weight"], state["linear1.bias"]))
    x = act(F.linear(x, state["linear2.weight"], state["linear2.bias"]))
    return F.linear(x, state["linear3.weight"], state["linear3.bias"])


# --------------------------------------------------------------------------------------------------
# DRMMTKS   (reranker/DRMMTKS.py:50-63)   -- SURVEY.md §8(f) rank 1
# --------------------------------------------------------------------------------------------------
def drmmtks_topk(table, doc, query, topk=10) -> torch.Tensor:
    """``torch.topk(cos_mat, k, dim=-1)`` of DRMMTKS.py:55-56 -> ``[B,Q,k]`` (descending; zeros of padded columns compete)."""
    sim = similarity_matrix(table, query, doc)
    top, _ = torch.topk(sim, k=topk, dim=-1)
    return top


def drmmtks_forward(state: dict, table, doc, query, query_idf, topk=10) -> torch.Tensor:
    """``DRMMTKS_class.forward(doc, query, query_idf)`` with ``gateType='IDF'`` (reranker/DRMMTKS.py:50-63) -> ``[B,1]``.

    (``gateType='TV'`` passes the int64 token ids to a Linear(E,1), DRMMTKS.py:59,43 -- it raises in the reference.)"""
    B, Q = query.shape
    q_mask = (query != 0).float()
    top = drmmtks_topk(table, doc, query, topk)
    z = torch.tanh(F.linear(top, state["ffw.0.weight"], state["ffw.0.bias"])).reshape(B, Q)
    gate = F.linear(query_idf.float()[:, :, None], state["gates.weight"]).reshape(B, Q) + (1 - q_mask) * -1e7
    w = F.softmax(gate, dim=1)
    x = (w * z).sum(dim=-1, keepdim=True)
    return F.linear(x, state["output_layer.weight"], state["output_layer.bias"])


# --------------------------------------------------------------------------------------------------
# ConvKNRM   (reranker/ConvKNRM.py:43-77, StackedSimilarityMatrix reranker/common.py:187-221)   -- SURVEY.md §8(f) rank 1
# --------------------------------------------------------------------------------------------------
def convknrm_ngram_reps(state: dict, table, toks, maxngram=3):
    """``conv[layer](pad(emb.permute(0,2,1))).permute(0,2,1)`` for n = 1..maxngram (ConvKNRM.py:47-50): list of ``[B,L,F]``.
    Conv1d(E, F, n) over the sequence zero-padded by n-1 on the right; no activation."""
    emb = F.embedding(toks, table).permute(0, 2, 1)
    reps = []
    for n in range(1, maxngram + 1):
        x = F.pad(emb, (0, n - 1), value=0.0) if n > 1 else emb
        reps.append(F.conv1d(x, state[f"convs.{n - 1}.0.weight"], state[f"convs.{n - 1}.0.bias"]).permute(0, 2, 1))
    return reps


def stacked_similarity(a, b, q_tok, d_tok, padding=0) -> torch.Tensor:
    """One cosine view of ``StackedSimilarityMatrix.forward`` (common.py:202-217) -> ``[B,1,Q,D]``."""
    a_den = a.norm(p=2, dim=2)[:, :, None] + 1e-9
    b_den = b.norm(p=2, dim=2)[:, None, :] + 1e-9
    sim = a.bmm(b.permute(0, 2, 1)) / (a_den * b_den)
    sim = torch.where((q_tok == padding)[:, :, None], torch.zeros_like(sim), sim)
    sim = torch.where((d_tok == padding)[:, None, :], torch.zeros_like(sim), sim)
    return sim[:, None]


def convknrm_features(state: dict, table, doc, query, maxngram=3, crossmatch=True, mus=None, sigmas=None) -> torch.Tensor:
    """The ``[B, K*VIEWS]`` tensor fed to ``combine`` (ConvKNRM.py:64-76); feature index = k * VIEWS + view,
    view = nq * maxngram + nd with crossmatch."""
    mus = KNRM_MUS if mus is None else mus
    sigmas = KNRM_SIGMAS if sigmas is None else sigmas
    a_reps = convknrm_ngram_reps(state, table, query, maxngram)
    b_reps = convknrm_ngram_reps(state, table, doc, maxngram)
    if crossmatch:
        views = [stacked_similarity(a, b, query, doc) for a in a_reps for b in b_reps]
    else:
        views = [stacked_similarity(a, b, query, doc) for a, b in zip(a_reps, b_reps)]
    simmats = torch.cat(views, dim=1)  # [B,VIEWS,Q,D]
    kernels = torch.stack([torch.exp(-0.5 * (simmats - m) * (simmats - m) / s / s) for m, s in zip(mus, sigmas)], dim=1)
    BATCH, KERNELS, VIEWS, QLEN, DLEN = kernels.shape
    kernels = kernels.reshape(BATCH, KERNELS * VIEWS, QLEN, DLEN)
    sim_rep = simmats.reshape(BATCH, 1, VIEWS, QLEN, DLEN).expand(BATCH, KERNELS, VIEWS, QLEN, DLEN).reshape(BATCH, KERNELS * VIEWS, QLEN, DLEN)
    result = kernels.sum(dim=3)
    mask = sim_rep.sum(dim=3) != 0.0
    result = torch.where(mask, (result + 1e-6).log(), mask.float())
    return result.sum(dim=2)


def convknrm_forward(state: dict, table, doc, query, query_idf=None, maxngram=3, crossmatch=True, singlefc=True,
                     scoretanh=False) -> torch.Tensor:
    """``ConvKNRM_class.forward(sentence, query_sentence, query_idf)`` (reranker/ConvKNRM.py:43-77) -> ``[B,1]``."""
    p = knrm_params_from_state(state)
    x = convknrm_features(state, table, doc, query, maxngram, crossmatch, p["mus"], p["sigmas"])
    if singlefc:
        x = F.linear(x, state["combine.0.weight"], state["combine.0.bias"])
    else:
        x = torch.tanh(F.linear(x, state["combine.0.weight"], state["combine.0.bias"]))
        x = F.linear(x, state["combine.2.weight"], state["combine.2.bias"])
    return torch.tanh(x) if scoretanh else x


# --------------------------------------------------------------------------------------------------
# losses   (reranker/common.py:7,96-103)
# --------------------------------------------------------------------------------------------------
def pair_hinge_loss(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    """MarginRankingLoss(margin=1, mean) with target +1: mean(max(0, 1 - (pos - neg)))."""
    return torch.clamp(1.0 - (pos - neg), min=0).mean()


def pair_softmax_loss(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    return torch.mean(1.0 - torch.stack([pos, neg], dim=1).softmax(dim=1)[:, 0])


# --------------------------------------------------------------------------------------------------
# monoBERT / BERT-MaxP   (reranker/ptBERTMaxP.py:52-96 + HF transformers BertForSequenceClassification)
# --------------------------------------------------------------------------------------------------
# The encoder arithmetic lives in a third-party dependency that is not under /root/reference:
# ``transformers~=4.9.2`` (reference setup.py:73).  The image has transformers 5.5.0 with the same
# BERT math (modeling_bert.py: embeddings -> 12 x {self-attention, output LN, erf-GELU FFN, output LN}
# -> pooler tanh -> classifier).  This is a restatement of that published algorithm; it is pinned
# against the installed HF implementation in tests/test_oracle.py.
def bert_layer(state: dict, p: str, x, num_heads: int, key_bias=None, eps: float = 1e-12):
    """One HF ``BertLayer`` (self-attention + output, intermediate + output) in eval mode; ``p`` = key prefix of the layer."""
    N, L, H = x.shape
    dh = H // num_heads
    g = lambda k: state[p + k]
    split = lambda t: t.reshape(N, L, num_heads, dh).transpose(1, 2)
    qh = split(F.linear(x, g("attention.self.query.weight"), g("attention.self.query.bias")))
    kh = split(F.linear(x, g("attention.self.key.weight"), g("attention.self.key.bias")))
    vh = split(F.linear(x, g("attention.self.value.weight"), g("attention.self.value.bias")))
    logits = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if key_bias is not None:
        logits = logits + key_bias
    att = torch.softmax(logits, dim=-1)
    ctx = (att @ vh).transpose(1, 2).reshape(N, L, H)
    y = F.linear(ctx, g("attention.output.dense.weight"), g("attention.output.dense.bias"))
    x = F.layer_norm(x + y, (H,), g("attention.output.LayerNorm.weight"), g("attention.output.LayerNorm.bias"), eps)
    y = F.gelu(F.linear(x, g("intermediate.dense.weight"), g("intermediate.dense.bias")))  # erf GELU
    y = F.linear(y, g("output.dense.weight"), g("output.dense.bias"))
    return F.layer_norm(x + y, (H,), g("output.LayerNorm.weight"), g("output.LayerNorm.bias"), eps)


def bert_hidden_states(state: dict, ids, mask, seg, num_heads: int, eps: float = 1e-12, prefix: str = "bert.") -> list:
    """HF BertModel ``hidden_states`` (embedding output + one entry per encoder layer), eval mode.

    ``state`` uses HF key names with ``prefix`` (``bert.`` for BertForSequenceClassification, empty for BertModel)."""
    N, L = ids.shape
    g = lambda k: state[prefix + k]
    x = F.embedding(ids, g("embeddings.word_embeddings.weight"))
    x = x + F.embedding(seg, g("embeddings.token_type_embeddings.weight"))
    x = x + g("embeddings.position_embeddings.weight")[:L][None]
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), g("embeddings.LayerNorm.weight"), g("embeddings.LayerNorm.bias"), eps)
    hidden = [x]
    dh = H // num_heads
    key_bias = torch.zeros(N, 1, 1, L, dtype=x.dtype)
    key_bias.masked_fill_(mask[:, None, None, :] == 0, torch.finfo(x.dtype).min)
    lp = prefix + "encoder.layer."
    n_layers = 1 + max(int(k[len(lp):].split(".")[0]) for k in state if k.startswith(lp))
    for i in range(n_layers):
        x = bert_layer(state, f"{prefix}encoder.layer.{i}.", x, num_heads, key_bias, eps)
        hidden.append(x)
    return hidden


def bert_logits(state: dict, ids, mask, seg, num_heads: int, eps: float = 1e-12) -> torch.Tensor:
    """``BertForSequenceClassification(ids, attention_mask=mask, token_type_ids=seg).logits`` in eval mode.

    ``state`` uses HF key names (``bert.embeddings.word_embeddings.weight`` ...); ids/mask/seg ``[N,L]`` int64."""
    x = bert_hidden_states(state, ids, mask, seg, num_heads, eps)[-1]
    pooled = torch.tanh(F.linear(x[:, 0], state["bert.pooler.dense.weight"], state["bert.pooler.dense.bias"]))
    return F.linear(pooled, state["classifier.weight"], state["classifier.bias"])


# --------------------------------------------------------------------------------------------------
# CEDR-KNRM   (reranker/CEDRKNRM.py:85-171)   -- SURVEY.md §8(f) rank 2
# --------------------------------------------------------------------------------------------------
def cedr_masked_simmat(emb, bert_mask, bert_seg, maxqlen_plus1: int):
    """``masked_simmats`` + ``_cos_simmat`` (CEDRKNRM.py:85-108) on ``emb [N, L-1, H]`` ([CLS] already removed)."""
    query_mask = bert_mask * (bert_seg == 0).to(emb.dtype)
    padded_query = (query_mask.unsqueeze(2) * emb)[:, :maxqlen_plus1]
    query_mask = query_mask[:, :maxqlen_plus1]
    doc_mask = bert_mask * (bert_seg == 1).to(emb.dtype)
    padded_doc = doc_mask.unsqueeze(2) * emb
    a_den = padded_query.norm(p=2, dim=2)[:, :, None] + 1e-9
    b_den = padded_doc.norm(p=2, dim=2)[:, None, :] + 1e-9
    sim = padded_query.bmm(padded_doc.permute(0, 2, 1)) / (a_den * b_den)
    sim = sim * query_mask[:, :, None] * doc_mask[:, None, :]
    return sim, doc_mask, query_mask


def cedr_knrm_features(hidden, bert_mask, bert_seg, batch_size, num_passages, maxqlen_plus1, mus, sigmas) -> torch.Tensor:
    """``CEDRKNRM_Class.knrm`` (CEDRKNRM.py:110-136) for one layer's hidden states ``[B*P, L, H]`` -> ``[B, K]``."""
    fm = bert_mask[:, 1:].to(hidden.dtype)
    sim, doc_mask, query_mask = cedr_masked_simmat(hidden[:, 1:], fm, bert_seg[:, 1:], maxqlen_plus1)
    Lm1 = sim.shape[2]
    sim = sim.view(batch_size, num_passages, maxqlen_plus1, Lm1)
    doc_mask = doc_mask.view(batch_size, num_passages, 1, Lm1)
    doc_simmat = torch.cat([sim[:, p] for p in range(num_passages)], dim=2)
    dmask = torch.cat([doc_mask[:, p] for p in range(num_passages)], dim=2)
    qmask = query_mask.view(batch_size, num_passages, -1, 1)[:, 0]
    kern = torch.stack([torch.exp(-0.5 * (doc_simmat - m) * (doc_simmat - m) / s / s) for m, s in zip(mus, sigmas)], dim=1)
    kern = kern * dmask.view(batch_size, 1, 1, -1) * qmask.view(batch_size, 1, -1, 1)
    feats = kern.sum(dim=3)
    feats = torch.log(torch.clamp(feats, min=1e-10)) * 0.01
    return feats.sum(dim=2)


def cedrknrm_forward(state: dict, bert_input, bert_mask, bert_seg, num_heads: int, maxqlen: int, simmat_layers, cls="avg",
                     combine_hidden=1024, eps: float = 1e-12) -> torch.Tensor:
    """``CEDRKNRM_Class.forward`` (CEDRKNRM.py:138-171) -> ``[B,1]``; inputs ``[B,P,L]`` int64; ``state`` = the module's state_dict
    (encoder under ``bert.``, ``kernels.kernels.{i}.mu|sigma``, ``combine.{0,1}``)."""
    B, P, L = bert_input.shape
    flat = lambda t: t.reshape(B * P, L)
    ids, mask, seg = flat(bert_input), flat(bert_mask), flat(bert_seg)
    hidden = bert_hidden_states(state, ids, mask, seg, num_heads, eps)
    kp = knrm_params_from_state(state)
    feats = []
    if cls:
        c = hidden[-1][:, 0, :].view(B, P, -1)
        feats.append(c.max(dim=1)[0] if cls == "max" else c.mean(dim=1))
    if -1 not in simmat_layers:
        feats += [cedr_knrm_features(hidden[l], mask, seg, B, P, maxqlen + 1, kp["mus"], kp["sigmas"]) for l in simmat_layers]
    x = torch.cat(feats, dim=1)
    x = F.linear(x, state["combine.0.weight"], state["combine.0.bias"])
    if combine_hidden:
        x = F.linear(x, state["combine.1.weight"], state["combine.1.bias"])
    return x


# --------------------------------------------------------------------------------------------------
# PARADE   (reranker/ptparade.py:55-78)   -- SURVEY.md §8(f) rank 2
# --------------------------------------------------------------------------------------------------
def parade_aggregate(state: dict, cls, batch_size: int, num_passages: int, num_heads: int, eps: float = 1e-12) -> torch.Tensor:
    """``aggregate_using_transformer`` (ptparade.py:55-68): ``cls [B*P, H]`` -> ``transformer_out_2[:, 0, :]`` ``[B, H]``."""
    H = cls.shape[-1]
    expanded = cls.view(batch_size, num_passages, H)
    tiled = state["initial_cls_embedding"].repeat(batch_size, 1).view(batch_size, 1, H)
    merged = torch.cat((tiled, expanded), dim=1) + state["full_position_embeddings"]
    out = bert_layer(state, "transformer_layer_1.", merged, num_heads, None, eps)
    out = bert_layer(state, "transformer_layer_2.", out, num_heads, None, eps)
    return out[:, 0, :]


def parade_forward(state: dict, doc_input, doc_mask, doc_seg, num_heads: int, eps: float = 1e-12) -> torch.Tensor:
    """``PTParade_Class.forward`` (ptparade.py:70-78) -> ``[B,1]``; inputs ``[B,P,L]`` int64; ``state`` = the module's state_dict."""
    B, P, L = doc_input.shape
    flat = lambda t: t.reshape(B * P, L)
    cls = bert_hidden_states(state, flat(doc_input), flat(doc_mask), flat(doc_seg), num_heads, eps)[-1][:, 0, :]
    agg = parade_aggregate(state, cls, B, P, num_heads, eps)
    return F.linear(agg, state["linear.weight"], state["linear.bias"])


def bert_maxp_aggregate(passage_scores, doc_mask, doc_seg, aggregation="max") -> torch.Tensor:
    """reranker/ptBERTMaxP.py:75-96.  ``passage_scores [B,P]``, ``doc_mask``/``doc_seg [B,P,L]`` -> ``[B]``.

    Note l.92: 'avg' divides by the sum of the passage mask over the WHOLE batch (a scalar)."""
    passage_mask = ((doc_mask * doc_seg).sum(dim=-1) > 5).long()
    if aggregation == "max":
        return passage_scores.max(dim=1)[0]
    if aggregation == "first":
        return passage_scores[:, 0]
    if aggregation == "sum":
        return torch.sum(passage_mask * passage_scores, dim=1)
    if aggregation == "avg":
        return torch.sum(passage_mask * passage_scores, dim=1) / torch.sum(passage_mask)
    raise ValueError("Unknown aggregation method: {}".format(aggregation))


def bert_maxp_forward(state: dict, doc_input, doc_mask, doc_seg, num_heads: int, aggregation="max") -> torch.Tensor:
    """``PTBERTMaxP_Class.predict_step`` (reranker/ptBERTMaxP.py:67-96): ``[B,P,L]`` int64 x3 -> ``[B]``."""
    B, P, L = doc_input.shape
    logits = bert_logits(state, doc_input.reshape(B * P, L), doc_mask.reshape(B * P, L), doc_seg.reshape(B * P, L), num_heads)
    return bert_maxp_aggregate(logits[:, 1].reshape(B, P), doc_mask, doc_seg, aggregation)

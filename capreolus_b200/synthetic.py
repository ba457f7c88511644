"""Seeded synthetic inputs in the layouts the reference extractors produce.

Layouts follow ``capreolus/extractor/embedtext.py:128-162`` (``query [B,Q] int64``, ``posdoc [B,D] int64``,
``query_idf [B,Q] f32``; id 0 = ``<pad>``, ids < 0 = out-of-vocabulary terms, ``reranker/common.py:174``)
and ``capreolus/extractor/bertpassage.py:268-346`` (``[CLS] q [SEP] passage [SEP] [PAD]...``; mask 1 on
real tokens; segment ids 0 for ``[CLS] q [SEP]`` and 1 from there to the end *including* the padding).

There is no network and no collection in this image, so every benchmark and parity test uses these.
"""
from __future__ import annotations

import numpy as np

VOCAB, EMB_DIM, MAXQLEN, MAXDOCLEN = 30000, 300, 32, 512
BERT_VOCAB, CLS, SEP = 30522, 101, 102


def embedding_table(vocab: int = VOCAB, dim: int = EMB_DIM, seed: int = 0) -> np.ndarray:
    """N(0,1) fp32 table with row 0 (``<pad>``) zero, as ``extractor/common.py:38-40`` builds it."""
    rng = np.random.default_rng(seed)
    table = rng.standard_normal((vocab, dim), dtype=np.float32)
    table[0] = 0.0
    return table


def _zipf_cdf(vocab: int) -> np.ndarray:
    w = 1.0 / np.arange(1, vocab, dtype=np.float64)
    return np.cumsum(w / w.sum())


def zipf_ids(rng: np.random.Generator, shape, vocab: int = VOCAB) -> np.ndarray:
    """Token ids in [1, vocab) with p(i) ~ 1/i, so repeated (exact-match) terms occur as in text."""
    u = rng.random(shape)
    ids = np.searchsorted(_zipf_cdf(vocab), u, side="left") + 1
    return np.minimum(ids, vocab - 1).astype(np.int64)


def parity_batch(batch: int = 64, qlen: int = MAXQLEN, dlen: int = MAXDOCLEN, vocab: int = VOCAB, seed: int = 1,
                 oov: bool = True, disjoint: bool = False) -> dict:
    """The PARITY set: ragged lengths right-padded with 0, shared OOV (negative) ids, one all-pad query.

    Returns numpy arrays ``query [B,Q] int64``, ``posdoc``/``negdoc [B,D] int64``, ``query_idf [B,Q] f32``.
    ``oov=False`` leaves out the negative ids (the reference DRMM indexes the table with the raw query ids,
    ``reranker/DRMM.py:109``, and raises IndexError on them).  ``disjoint=True`` draws query terms from the even
    and document terms from the odd ids, so that no in-vocabulary exact match occurs (the reference's fp32 result for
    identical tokens is rounding noise -- DESIGN.md "Exact matches" -- and some checks need inputs free of it).
    """
    rng = np.random.default_rng(seed)
    query = zipf_ids(rng, (batch, qlen), vocab)
    docs = [zipf_ids(rng, (batch, dlen), vocab) for _ in range(2)]
    qn = rng.integers(1, qlen + 1, size=batch)
    dn = [rng.integers(min(16, dlen), dlen + 1, size=batch) for _ in range(2)]
    if batch >= 4:
        qn[0], dn[0][0], dn[1][0] = qlen, dlen, dlen  # one completely full pair
        qn[1] = 1  # a one-term query
        dn[0][2] = min(16, dlen)  # shortest document
    if disjoint:
        half = (vocab - 1) // 2
        query = 2 * np.minimum(zipf_ids(rng, (batch, qlen), half + 1), half)  # even ids in [2, vocab)
        docs = [2 * np.minimum(zipf_ids(rng, (batch, dlen), half + 1), half) - 1 for _ in range(2)]  # odd ids in [1, vocab)
        oov = False
    # every query term also occurs somewhere in its documents (soft-TF needs exact matches to be exercised)
    for b in range(batch if not disjoint else 0):
        for doc in docs:
            pos = rng.integers(0, dlen, size=qlen // 4 + 1)
            doc[b, pos] = query[b, rng.integers(0, qlen, size=pos.shape[0])]
    # out-of-vocabulary terms: negative ids; identical negatives in query and doc are exact matches
    n_oov_rows = max(1, batch // 8) if oov else 0
    for b in rng.choice(batch, size=n_oov_rows, replace=False) if oov else []:
        oov_ids = -rng.integers(1, 50, size=3)
        query[b, rng.integers(0, max(1, qn[b]), size=3)] = oov_ids
        for doc, n in zip(docs, dn):
            doc[b, rng.integers(0, n[b], size=2)] = oov_ids[:2]  # oov_ids[2] has no partner in the doc
    ar_q, ar_d = np.arange(qlen)[None, :], np.arange(dlen)[None, :]
    query[ar_q >= qn[:, None]] = 0
    for doc, n in zip(docs, dn):
        doc[ar_d >= n[:, None]] = 0
    if batch >= 4:
        query[3, :] = 0  # an all-pad query row (reference quirk: contributes 0, not log(eps))
    idf = rng.random((batch, qlen), dtype=np.float32) * 8.0
    idf[query == 0] = 0.0
    if batch >= 6:
        idf[5, :] = 0.0  # the real EmbedText pipeline feeds all-zero idf (embedtext.py:86-96)
    return {"query": query, "posdoc": docs[0], "negdoc": docs[1], "query_idf": idf}


def throughput_batch(n: int, qlen: int = MAXQLEN, dlen: int = MAXDOCLEN, vocab: int = VOCAB, seed: int = 2) -> dict:
    """The THROUGHPUT set: full-length (no padding) zipf ids, int64 like the reference (``np.long``)."""
    rng = np.random.default_rng(seed)
    return {
        "query": zipf_ids(rng, (n, qlen), vocab),
        "posdoc": zipf_ids(rng, (n, dlen), vocab),
        "query_idf": (rng.random((n, qlen), dtype=np.float32) * 8.0),
    }


def train_triples(n: int, qlen: int = MAXQLEN, dlen: int = MAXDOCLEN, vocab: int = VOCAB, seed: int = 4,
                  disjoint: bool = False) -> dict:
    """The TRAIN set: (query, posdoc, negdoc) triples; posdoc shares more terms with the query than negdoc
    (``disjoint=True``: no shared terms at all, see ``parity_batch``)."""
    rng = np.random.default_rng(seed)
    out = parity_batch(n, qlen, dlen, vocab, seed=seed, disjoint=disjoint)
    pos = out["posdoc"]
    for b in range(n if not disjoint else 0):
        real_q = out["query"][b][out["query"][b] > 0]
        nd = int((pos[b] != 0).sum())
        if real_q.size and nd:
            where = rng.integers(0, nd, size=8)
            pos[b, where] = real_q[rng.integers(0, real_q.size, size=8)]
    return out


def bert_batch(n: int, seqlen: int = 512, qlen: int = 32, vocab: int = BERT_VOCAB, seed: int = 3, numpassages: int = 1,
               ragged: bool = True) -> dict:
    """The BERT set: ``[CLS] q(qlen) [SEP] d [SEP]`` padded to ``seqlen``; arrays are ``[n, P, L] int64``."""
    rng = np.random.default_rng(seed)
    lo = min(1000, vocab // 2)
    if not ragged:  # full-length sequences (throughput runs): vectorised, 125 000 sequences in a second instead of a Python loop
        ids = rng.integers(lo, vocab, size=(n, numpassages, seqlen), dtype=np.int64)
        ids[..., 0], ids[..., qlen + 1], ids[..., seqlen - 1] = CLS, SEP, SEP
        mask = np.ones_like(ids)
        seg = np.zeros_like(ids)
        seg[..., qlen + 2:] = 1
        return {"pos_bert_input": ids, "pos_mask": mask, "pos_seg": seg}
    ids = np.zeros((n, numpassages, seqlen), dtype=np.int64)
    mask = np.zeros_like(ids)
    seg = np.zeros_like(ids)
    for b in range(n):
        for p in range(numpassages):
            max_d = seqlen - qlen - 3
            nd = int(rng.integers(max(1, max_d // 4), max_d + 1)) if ragged and (b + p) % 4 else max_d
            if ragged and numpassages > 1 and p == numpassages - 1 and b % 3 == 0:
                nd = 2  # a nearly empty trailing passage (passage_mask = 0 for sum/avg pooling)
            toks = [CLS] + list(rng.integers(lo, vocab, size=qlen)) + [SEP] + list(rng.integers(lo, vocab, size=nd)) + [SEP]
            ids[b, p, : len(toks)] = toks
            mask[b, p, : len(toks)] = 1
            seg[b, p, qlen + 2:] = 1  # stays 1 through the padding (tests/test_extractor.py:708-713)
    return {"pos_bert_input": ids, "pos_mask": mask, "pos_seg": seg}


def cedr_batch(n: int, numpassages: int, seqlen: int, maxqlen: int, vocab: int = BERT_VOCAB, seed: int = 6) -> dict:
    """Passage inputs in the ``pooledbertpassage`` layout CEDR-KNRM consumes (``capreolus/extractor/bertpassage.py:268-284``):
    ``[CLS] q [SEP] passage [SEP] [PAD]...`` with the SAME query in every passage of a document, ragged query lengths
    (1..maxqlen), ragged passages and some completely empty trailing passages; arrays are ``[n, P, L] int64``."""
    rng = np.random.default_rng(seed)
    ids = np.zeros((n, numpassages, seqlen), dtype=np.int64)
    mask = np.zeros_like(ids)
    seg = np.zeros_like(ids)
    lo = min(1000, vocab // 2)
    for b in range(n):
        ql = maxqlen if b == 0 else int(rng.integers(1, maxqlen + 1))
        q = list(rng.integers(lo, vocab, size=ql))
        for p in range(numpassages):
            max_d = seqlen - ql - 3
            nd = max_d if (b + p) % 3 == 0 else int(rng.integers(1, max_d + 1))
            if numpassages > 1 and p == numpassages - 1 and b % 2 == 1:
                nd = 0  # an empty trailing passage: "[CLS] q [SEP] [SEP]"
            toks = [CLS] + q + [SEP] + list(rng.integers(lo, vocab, size=nd)) + [SEP]
            ids[b, p, : len(toks)] = toks
            mask[b, p, : len(toks)] = 1
            seg[b, p, ql + 2:] = 1
    return {"pos_bert_input": ids, "pos_mask": mask, "pos_seg": seg}

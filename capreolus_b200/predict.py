"""Host-side predict loop: host batches in, host scores out.

Mirrors the hot loop of ``PytorchTrainer.predict`` (``capreolus/trainer/pytorch.py:338-348``): move the batch to
the device, ``reranker.test(batch)``, bring the scores back.  The reference does this synchronously per batch
(``.cpu()`` every iteration); here host->device copies of chunk i+1 run on a copy stream while chunk i is being
scored, and the scores of all chunks come back with one synchronisation at the end.
"""
from __future__ import annotations

import torch


def _narrowest(t: torch.Tensor) -> torch.dtype:
    """The narrowest signed integer dtype that holds every value of an int64 tensor (ids: 0 = <pad>, < 0 = OOV keep their sign)."""
    if t.numel() == 0:
        return torch.int16
    lo, hi = int(t.min()), int(t.max())
    if -32768 <= lo and hi <= 32767:
        return torch.int16
    if -(2 ** 31) <= lo and hi <= 2 ** 31 - 1:
        return torch.int32
    return torch.int64


class PinnedBatch:
    """A host batch (dict of tensors / lists as the DataLoader collates them) moved to pinned memory once.

    ``narrow_ids=True`` (default) stores int64 tensors (token ids, BERT masks / segment ids) in the narrowest of int16 / int32 /
    int64 that holds their values -- the reference extractor emits ``np.long`` (``embedtext.py:146-147``), 4 352 B per
    (|q|=32, |d|=512) pair, which bounds the host -> device leg of the predict loop; a 30 k vocabulary fits int16 (1 088 B).
    ``PipelinedPredictor`` widens them on the device (``capr_widen_ids``), so the rerankers still see int64.  ``wide`` names
    the tensors that were narrowed."""

    def __init__(self, batch: dict, narrow_ids: bool = True):
        self.tensors, self.wide = {}, {}
        for k, v in batch.items():
            if not torch.is_tensor(v):
                continue
            if narrow_ids and v.dtype == torch.int64:
                dt = _narrowest(v)
                if dt != torch.int64:
                    self.wide[k] = torch.int64
                    v = v.to(dt)
            self.tensors[k] = v if v.is_pinned() else v.contiguous().pin_memory()
        self.n = len(next(iter(self.tensors.values())))
        self.bytes_per_item = sum(v[0].numel() * v.element_size() for v in self.tensors.values()) if self.n else 0


class PipelinedPredictor:
    """Double-buffered H2D / score / D2H pipeline around ``reranker.test``.

    ``ramp``: the first copy cannot overlap anything, so the schedule starts with chunks of ``chunk/8, chunk/4, chunk/2``
    items (the exposed copy shrinks 8x) and then continues with full chunks."""

    def __init__(self, reranker, device, chunk: int = 16384, ramp: bool = True):
        self.reranker, self.device, self.chunk, self.ramp = reranker, torch.device(device), int(chunk), bool(ramp)
        if self.chunk <= 0:
            raise ValueError("PipelinedPredictor: chunk must be positive")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = None
        self._out_host = None

    def _buffers(self, pb: PinnedBatch):
        sig = tuple((k, v.dtype, tuple(v.shape[1:]), k in pb.wide) for k, v in pb.tensors.items())
        if self._bufs is None or self._bufs[0] != sig:
            mk = lambda v, dt: torch.empty((self.chunk,) + tuple(v.shape[1:]), dtype=dt, device=self.device)
            staged = [{k: mk(v, v.dtype) for k, v in pb.tensors.items()} for _ in range(2)]
            wide = [{k: mk(pb.tensors[k], dt) for k, dt in pb.wide.items()} for _ in range(2)]  # int64 rows the rerankers consume
            self._bufs = (sig, staged, wide)
        if self._out_host is None or self._out_host.shape[0] < pb.n:
            self._out_host = torch.empty(pb.n, dtype=torch.float32).pin_memory()
        return self._bufs[1], self._bufs[2]

    def schedule(self, n: int) -> list:
        """``[(lo, hi), ...]`` covering ``range(n)``: ramp-up chunks first (if enabled), then full chunks."""
        sizes, lo, out = ([self.chunk // 8, self.chunk // 4, self.chunk // 2] if self.ramp else []), 0, []
        for sz in sizes:
            if sz > 0 and lo + sz < n:
                out.append((lo, lo + sz))
                lo += sz
        while lo < n:
            out.append((lo, min(n, lo + self.chunk)))
            lo += self.chunk
        return out

    @torch.no_grad()
    def predict(self, pb: PinnedBatch) -> torch.Tensor:
        """Scores for every item of ``pb`` as a host fp32 tensor ``[n]`` (a fresh tensor: a later call does not overwrite it)."""
        from capreolus_b200 import _lib

        with torch.cuda.device(self.device):
            staged, wide = self._buffers(pb)
            main = torch.cuda.current_stream(self.device)
            spans = self.schedule(pb.n)
            copied = [torch.cuda.Event() for _ in spans]
            scored = [torch.cuda.Event() for _ in spans]
            self.copy_stream.wait_stream(main)
            for i, (lo, hi) in enumerate(spans):
                buf, wbuf = staged[i % 2], wide[i % 2]
                with torch.cuda.stream(self.copy_stream):
                    if i >= 2:
                        self.copy_stream.wait_event(scored[i - 2])  # buffer reuse
                    for k, v in pb.tensors.items():
                        buf[k][: hi - lo].copy_(v[lo:hi], non_blocking=True)
                    copied[i].record(self.copy_stream)
                main.wait_event(copied[i])
                view = {}
                for k, v in buf.items():
                    if k in wbuf:  # narrow ids -> int64 on the device (2 + 8 bytes of HBM traffic per token)
                        src, dst = v[: hi - lo], wbuf[k][: hi - lo]
                        _lib.check(_lib.lib().capr_widen_ids(src.data_ptr(), src.element_size(), src.numel(), dst.data_ptr(),
                                                            main.cuda_stream))
                        view[k] = dst
                    else:
                        view[k] = v[: hi - lo]
                scores = self.reranker.test(view)
                self._out_host[lo:hi].copy_(scores.view(-1), non_blocking=True)
                scored[i].record(main)
            main.synchronize()
        return self._out_host[: pb.n].clone()


# ---------------------------------------------------------------------------------------------------------------------
# packed id store -> on-device batch assembly -> scoring -> on-device ranking -> TREC run
# ---------------------------------------------------------------------------------------------------------------------
class PackedIdStore:
    """A tokenised collection (queries or documents) as ONE flat int32 id array + int64 offsets.

    Replaces the per-item Python lists behind ``EmbedText.id2vec`` (``capreolus/extractor/embedtext.py:128-151``:
    ``qid2toks`` / ``get_doc_tokens`` -> ``_tok2vec``).  Items keep their string ids (``names``); ``idf`` optionally holds
    one fp32 per token (the ``idfs`` of ``id2vec``)."""

    def __init__(self, names, flat, offsets, idf=None):
        import numpy as np

        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        if len(self.index) != len(self.names):
            raise ValueError("PackedIdStore: duplicate item ids")
        self.flat = torch.as_tensor(np.ascontiguousarray(flat, dtype=np.int32))
        self.offsets = torch.as_tensor(np.ascontiguousarray(offsets, dtype=np.int64))
        if self.offsets.numel() != len(self.names) + 1 or int(self.offsets[0]) != 0 or int(self.offsets[-1]) != self.flat.numel():
            raise ValueError("PackedIdStore: offsets must have len(names)+1 entries, start at 0 and end at len(flat)")
        if self.offsets.numel() > 1 and bool((self.offsets[1:] < self.offsets[:-1]).any()):
            raise ValueError("PackedIdStore: offsets must be non-decreasing")
        self.idf = None if idf is None else torch.as_tensor(np.ascontiguousarray(idf, dtype=np.float32))
        if self.idf is not None and self.idf.numel() != self.flat.numel():
            raise ValueError("PackedIdStore: idf must be parallel to the flat id array")

    @classmethod
    def from_lists(cls, items: dict, idf: dict | None = None):
        """``items``: ``{id: [token ids]}`` (insertion order is kept); ``idf``: ``{id: [fp32 per token]}``."""
        import numpy as np

        names = list(items)
        lens = np.fromiter((len(items[n]) for n in names), dtype=np.int64, count=len(names))
        offsets = np.zeros(len(names) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        flat = np.concatenate([np.asarray(items[n], dtype=np.int32) for n in names]) if names else np.zeros(0, np.int32)
        idfs = None
        if idf is not None:
            idfs = np.concatenate([np.asarray(idf[n], dtype=np.float32) for n in names]) if names else np.zeros(0, np.float32)
        return cls(names, flat, offsets, idfs)

    def __len__(self):
        return len(self.names)

    def to(self, device):
        self.flat, self.offsets = self.flat.to(device), self.offsets.to(device)
        if self.idf is not None:
            self.idf = self.idf.to(device)
        return self


class PairAssembler:
    """``capr_assemble_pairs``: (query index, doc index) vectors -> the padded int64 batch the rerankers consume."""

    def __init__(self, queries: PackedIdStore, docs: PackedIdStore, maxqlen: int, maxdoclen: int, device):
        self.device = torch.device(device)
        self.queries, self.docs = queries.to(self.device), docs.to(self.device)
        self.Q, self.D = int(maxqlen), int(maxdoclen)

    def assemble(self, qidx: torch.Tensor, didx: torch.Tensor, out: dict | None = None, rows: int | None = None) -> dict:
        """Fill ``out`` (allocated when ``None``; ``rows`` = capacity to allocate, default ``len(qidx)``) with the padded rows."""
        from capreolus_b200 import _lib

        _lib.require_cuda(qidx, didx)
        if qidx.dtype != torch.int32 or didx.dtype != torch.int32:
            raise ValueError("PairAssembler.assemble: index vectors must be int32")
        n = qidx.shape[0]
        if out is None:
            cap = n if rows is None else int(rows)
            out = {"query": torch.empty((cap, self.Q), dtype=torch.int64, device=self.device),
                   "posdoc": torch.empty((cap, self.D), dtype=torch.int64, device=self.device),
                   "query_idf": torch.empty((cap, self.Q), dtype=torch.float32, device=self.device)}
        if n == 0:
            return out
        q, d = self.queries, self.docs
        _lib.check(_lib.lib().capr_assemble_pairs(
            q.flat.data_ptr(), q.offsets.data_ptr(), len(q), d.flat.data_ptr(), d.offsets.data_ptr(), len(d), _lib.ptr(q.idf),
            qidx.contiguous().data_ptr(), didx.contiguous().data_ptr(), n, self.Q, self.D, out["query"].data_ptr(), out["posdoc"].data_ptr(),
            out["query_idf"].data_ptr(), _lib.current_stream(self.device)))
        return out


class BertPairAssembler:
    """``capr_assemble_bert_pairs``: (query index, doc index) vectors -> ``pos_bert_input / pos_mask / pos_seg [N,P,L]`` int64, the
    rows ``BertPassage.id2vec`` builds per pair (``capreolus/extractor/bertpassage.py:203-232, 268-346``; sliding-window
    passages).  The stores hold WordPiece ids (the tokenizer stays on the host, once per collection)."""

    def __init__(self, queries: PackedIdStore, docs: PackedIdStore, maxqlen: int, maxseqlen: int, numpassages: int, passagelen: int, stride: int,
                 device, padq: bool = False, cls_id: int = 101, sep_id: int = 102, pad_id: int = 0):
        self.device = torch.device(device)
        self.queries, self.docs = queries.to(self.device), docs.to(self.device)
        self.Q, self.L, self.P = int(maxqlen), int(maxseqlen), int(numpassages)
        self.passagelen, self.stride, self.padq = int(passagelen), int(stride), bool(padq)
        self.cls_id, self.sep_id, self.pad_id = int(cls_id), int(sep_id), int(pad_id)

    def assemble(self, qidx: torch.Tensor, didx: torch.Tensor, out: dict | None = None, rows: int | None = None) -> dict:
        """Fill ``out`` (allocated when ``None``; ``rows`` = capacity to allocate, default ``len(qidx)``) with the passage rows."""
        from capreolus_b200 import _lib

        _lib.require_cuda(qidx, didx)
        if qidx.dtype != torch.int32 or didx.dtype != torch.int32:
            raise ValueError("BertPairAssembler.assemble: index vectors must be int32")
        n = qidx.shape[0]
        if out is None:
            cap = n if rows is None else int(rows)
            out = {k: torch.empty((cap, self.P, self.L), dtype=torch.int64, device=self.device) for k in ("pos_bert_input", "pos_mask", "pos_seg")}
        if n == 0:
            return out
        q, d = self.queries, self.docs
        _lib.check(_lib.lib().capr_assemble_bert_pairs(
            q.flat.data_ptr(), q.offsets.data_ptr(), len(q), d.flat.data_ptr(), d.offsets.data_ptr(), len(d), qidx.contiguous().data_ptr(),
            didx.contiguous().data_ptr(), n, self.P, self.L, self.Q, int(self.padq), self.passagelen, self.stride, self.cls_id, self.sep_id,
            self.pad_id, out["pos_bert_input"].data_ptr(), out["pos_mask"].data_ptr(), out["pos_seg"].data_ptr(), _lib.current_stream(self.device)))
        return out


def rank_by_query(scores: torch.Tensor, seg_off: torch.Tensor, max_segment: int):
    """``capr_rank_by_query``: (float16-rounded scores ``[N]`` fp32, per-query order ``[N]`` int32); see include/capr_b200.h."""
    from capreolus_b200 import _lib

    _lib.require_cuda(scores, seg_off)
    scores = scores.contiguous().float()
    rounded = torch.empty_like(scores)
    order = torch.empty(scores.shape[0], dtype=torch.int32, device=scores.device)
    _lib.check(_lib.lib().capr_rank_by_query(scores.data_ptr(), seg_off.contiguous().data_ptr(), seg_off.shape[0] - 1, int(max_segment),
                                            rounded.data_ptr(), order.data_ptr(), _lib.current_stream(scores.device)))
    return rounded, order


def write_trec_run(preds: dict, outfn, mode="wt"):
    """``Searcher.write_trec_run`` (capreolus/searcher/__init__.py:48-58), same file format and ordering."""
    with open(outfn, mode) as outf:
        for qid in sorted(preds.keys(), key=lambda k: int(k)):
            for rank, (docid, score) in enumerate(sorted(preds[qid].items(), key=lambda x: x[1], reverse=True), start=1):
                print(f"{qid} Q0 {docid} {rank} {score} capreolus", file=outf)


class RunPredictor:
    """``PytorchTrainer.predict`` (capreolus/trainer/pytorch.py:310-353) for a packed collection.

    ``predict(reranker, qid_to_docids, pred_fn)`` scores every (qid, docid) candidate, returns the reference's
    ``{qid: {docid: float16-rounded score}}`` dict and (optionally) writes the TREC run.  Per chunk only the two int32 index
    vectors cross PCIe (8 B per pair instead of the 4 352 B of padded int64 ids); rows are assembled, scored, rounded and
    ranked on the device, and the run file is written from the device-computed order (no Python sort)."""

    def __init__(self, assembler: PairAssembler, chunk: int = 16384):
        self.assembler, self.chunk = assembler, int(chunk)
        self.device = assembler.device
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = None

    @torch.no_grad()
    def score_indices(self, reranker, qidx_host: torch.Tensor, didx_host: torch.Tensor) -> torch.Tensor:
        """Device scores ``[N]`` for pinned host int32 index vectors (H2D of chunk i+1 overlaps scoring of chunk i)."""
        n = qidx_host.shape[0]
        main = torch.cuda.current_stream(self.device)
        scores = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_stream(main)
            qd = qidx_host.to(self.device, non_blocking=True)
            dd = didx_host.to(self.device, non_blocking=True)
        main.wait_stream(self.copy_stream)
        qd.record_stream(main), dd.record_stream(main)
        # one set of assembly buffers of the full chunk size, reused by every chunk and every call (a ragged last chunk takes
        # a leading view: no allocator traffic inside the loop)
        if self._bufs is None:
            self._bufs = self.assembler.assemble(qd[:0], dd[:0], None, rows=self.chunk)
        for lo in range(0, n, self.chunk):
            hi = min(n, lo + self.chunk)
            view = {k: v[: hi - lo] for k, v in self._bufs.items()}
            self.assembler.assemble(qd[lo:hi], dd[lo:hi], view)
            scores[lo:hi] = reranker.test(view).view(-1)
        return scores

    @torch.no_grad()
    def predict(self, reranker, qid_to_docids: dict, pred_fn=None) -> dict:
        import os

        import numpy as np

        a = self.assembler
        model = reranker.model.to(self.device)
        model.eval()
        qids = list(qid_to_docids)
        lens = np.fromiter((len(qid_to_docids[q]) for q in qids), dtype=np.int64, count=len(qids))
        seg = np.zeros(len(qids) + 1, dtype=np.int64)
        np.cumsum(lens, out=seg[1:])
        n = int(seg[-1])
        try:
            qidx = np.repeat(np.fromiter((a.queries.index[q] for q in qids), dtype=np.int32, count=len(qids)), lens)
            didx = np.fromiter((a.docs.index[d] for q in qids for d in qid_to_docids[q]), dtype=np.int32, count=n)
        except KeyError as e:  # the reference raises MissingDocError while predicting (sampler/__init__.py:229-232)
            raise KeyError(f"got none features for prediction: unknown query or document id {e}") from None
        scores = self.score_indices(reranker, torch.from_numpy(qidx).pin_memory(), torch.from_numpy(didx).pin_memory())
        rounded, order = rank_by_query(scores, torch.from_numpy(seg).to(self.device), int(lens.max()) if len(lens) else 0)
        rounded, order = rounded.cpu().numpy(), order.cpu().numpy()
        preds = {}
        for i, q in enumerate(qids):
            docs, s = qid_to_docids[q], rounded[seg[i]:seg[i + 1]]
            preds[q] = {d: float(x) for d, x in zip(docs, s)}
        if pred_fn is not None:
            os.makedirs(os.path.dirname(os.path.abspath(pred_fn)), exist_ok=True)
            with open(pred_fn, "wt") as outf:
                for i in sorted(range(len(qids)), key=lambda j: int(qids[j])):
                    q, docs, base = qids[i], qid_to_docids[qids[i]], int(seg[i])
                    for rank in range(int(lens[i])):
                        pos = int(order[base + rank])
                        print(f"{q} Q0 {docs[pos]} {rank + 1} {float(rounded[base + pos])} capreolus", file=outf)
        return preds

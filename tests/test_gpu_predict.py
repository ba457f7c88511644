"""GPU: the host-side predict pipeline (H2D / score / D2H overlap) returns exactly what reranker.test returns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Extractor:
    def __init__(self, table, Q, D):
        self.embeddings = table
        self.config = {"maxqlen": Q, "maxdoclen": D}


@pytest.mark.parametrize("n,chunk", [(1000, 256), (513, 512), (7, 64), (2048, 2048)])
def test_pipelined_predict_equals_direct_test(n, chunk):
    from capreolus_b200 import reranker as R, synthetic
    from capreolus_b200.predict import PinnedBatch, PipelinedPredictor

    Q, D, V, E = 32, 512, 3000, 300
    rr = R.KNRM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    rr.build_model().to(DEV).eval()
    host = {k: torch.from_numpy(v) for k, v in synthetic.throughput_batch(n, Q, D, V, seed=9).items()}
    with torch.no_grad():
        direct = rr.test({k: v.to(DEV) for k, v in host.items()}).cpu()
    # ids travel as int16 when the vocabulary fits (default), as the reference's int64 with narrow_ids=False
    for narrow, per_item in ((True, (Q + D) * 2 + Q * 4), (False, (Q + D) * 8 + Q * 4)):
        pb = PinnedBatch(host, narrow_ids=narrow)
        assert pb.n == n and pb.bytes_per_item == per_item
        assert (pb.tensors["posdoc"].dtype == torch.int16) == narrow
        pred = PipelinedPredictor(rr, DEV, chunk=chunk)
        first = None
        for _ in range(2):  # buffers are reused across calls
            out = pred.predict(pb)
            assert out.shape == (n,)
            assert torch.equal(out.cpu(), direct)
            if first is None:
                first = out
        first_copy = first.clone()
        pred.predict(PinnedBatch({k: v.flip(0) for k, v in host.items()}))  # a later call must not overwrite earlier results
        assert torch.equal(first, first_copy)


def test_narrow_ids_keep_oov_signs_and_wide_vocabularies():
    """capr_widen_ids sign-extends: negative (OOV) ids survive the int16 / int32 trip; ids beyond int16 select int32."""
    from capreolus_b200 import reranker as R, synthetic
    from capreolus_b200.predict import PinnedBatch, PipelinedPredictor

    Q, D, V, E = 8, 40, 40000, 32
    rr = R.KNRM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    rr.build_model().to(DEV).eval()
    small = {k: torch.from_numpy(v) for k, v in synthetic.parity_batch(12, Q, D, 3000, seed=5, oov=True).items()}  # has negative ids
    assert int(small["query"].min()) < 0
    wide = {k: v.clone() for k, v in small.items()}
    wide["posdoc"][:, 0] = 39999  # does not fit int16
    for host, want in ((small, torch.int16), (wide, torch.int32)):
        pb = PinnedBatch(host)
        assert pb.tensors["posdoc"].dtype == want and pb.tensors["query_idf"].dtype == torch.float32
        with torch.no_grad():
            direct = rr.test({k: v.to(DEV) for k, v in host.items()}).cpu()
        assert torch.equal(PipelinedPredictor(rr, DEV, chunk=5).predict(pb).cpu(), direct)


def test_table_is_rebuilt_when_the_embedding_changes():
    """PreparedTable caches derived data keyed on the weight's version counter (finetune=True steps, load_state_dict)."""
    from capreolus_b200 import reranker as R, synthetic

    Q, D, V, E = 8, 40, 500, 50
    rr = R.KNRM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    model = rr.build_model().to(DEV).eval()
    b = {k: torch.from_numpy(v).to(DEV) for k, v in synthetic.parity_batch(6, Q, D, V, seed=3).items()}
    with torch.no_grad():
        s0 = rr.test(b).clone()
        model.embedding.weight.mul_(-1.0)  # cosines are invariant to a global sign flip
        s1 = rr.test(b).clone()
        model.embedding.weight[1:50].normal_()
        s2 = rr.test(b)
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-5)
    assert not torch.allclose(s0, s2, rtol=1e-3, atol=1e-3)


# ---- on-device batch assembly, ranking, run files ------------------------------------------------
def _collection(n_q, n_d, Q, D, V, seed):
    rng = np.random.default_rng(seed)
    queries = {str(100 + i): rng.integers(1, V, size=rng.integers(0, Q + 5)).tolist() for i in range(n_q)}
    docs = {f"doc{i}": rng.integers(1, V, size=rng.integers(0, D + 40)).tolist() for i in range(n_d)}
    idf = {k: rng.random(len(v)).astype(np.float32).tolist() for k, v in queries.items()}
    return queries, docs, idf


def _padlist(x, n, pad=0):  # capreolus/utils/common.py:99-111
    x = list(x)[:n]
    return x + [pad] * (n - len(x))


def test_assemble_pairs_equals_padlist():
    from capreolus_b200.predict import PackedIdStore, PairAssembler

    Q, D, V = 8, 50, 1000
    queries, docs, idf = _collection(7, 40, Q, D, V, seed=1)
    asm = PairAssembler(PackedIdStore.from_lists(queries, idf), PackedIdStore.from_lists(docs), Q, D, DEV)
    rng = np.random.default_rng(2)
    qi = rng.integers(0, 7, size=333).astype(np.int32)
    di = rng.integers(0, 40, size=333).astype(np.int32)
    out = asm.assemble(torch.from_numpy(qi).to(DEV), torch.from_numpy(di).to(DEV))
    qn, dn = list(queries), list(docs)
    want_q = np.array([_padlist(queries[qn[i]], Q) for i in qi], dtype=np.int64)
    want_d = np.array([_padlist(docs[dn[i]], D) for i in di], dtype=np.int64)
    want_idf = np.array([_padlist(idf[qn[i]], Q, 0.0) for i in qi], dtype=np.float32)
    assert out["query"].dtype == torch.int64 and np.array_equal(out["query"].cpu().numpy(), want_q)
    assert np.array_equal(out["posdoc"].cpu().numpy(), want_d)
    assert np.array_equal(out["query_idf"].cpu().numpy(), want_idf)
    # out-of-range indices give all-pad rows; an empty batch is a no-op
    bad = asm.assemble(torch.tensor([-1, 99], dtype=torch.int32, device=DEV), torch.tensor([400, -5], dtype=torch.int32, device=DEV))
    assert int(bad["query"].abs().sum()) == 0 and int(bad["posdoc"].abs().sum()) == 0
    assert asm.assemble(torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV))["query"].shape == (0, Q)


@pytest.mark.parametrize("sizes", [[1, 5, 128, 0, 77], [1000, 3, 1024], [4096, 1500, 2]])
def test_rank_by_query_equals_python_stable_sort(sizes):
    from capreolus_b200.predict import rank_by_query

    rng = np.random.default_rng(sum(sizes))
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    scores = (rng.standard_normal(int(seg[-1])) * 3).astype(np.float32)
    scores[rng.integers(0, scores.size, size=scores.size // 3)] = 0.5  # plenty of ties
    rounded, order = rank_by_query(torch.from_numpy(scores).to(DEV), torch.from_numpy(seg).to(DEV), max(sizes))
    rounded, order = rounded.cpu().numpy(), order.cpu().numpy()
    want_r = scores.astype(np.float16).astype(np.float32)  # trainer/pytorch.py:347
    assert np.array_equal(rounded, want_r)
    for i in range(len(sizes)):
        s = want_r[seg[i]:seg[i + 1]]
        want = [j for j, _ in sorted(enumerate(s.tolist()), key=lambda x: x[1], reverse=True)]  # searcher/__init__.py:54
        assert order[seg[i]:seg[i + 1]].tolist() == want


def test_run_predictor_matches_reference_predict_loop(tmp_path):
    """RunPredictor.predict == the reference loop: id2vec rows -> reranker.test -> float16 -> {qid: {docid: score}} -> write_trec_run."""
    from capreolus_b200 import reranker as R, synthetic
    from capreolus_b200.predict import PackedIdStore, PairAssembler, RunPredictor, write_trec_run

    Q, D, V, E = 8, 50, 1000, 50
    queries, docs, idf = _collection(9, 120, Q, D, V, seed=5)
    rr = R.DRMM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    rr.build_model().to(DEV).eval()
    rng = np.random.default_rng(6)
    dn = list(docs)
    cands = {q: [dn[j] for j in rng.choice(len(dn), size=rng.integers(1, 60), replace=False)] for q in queries}
    asm = PairAssembler(PackedIdStore.from_lists(queries, idf), PackedIdStore.from_lists(docs), Q, D, DEV)
    preds = RunPredictor(asm, chunk=100).predict(rr, cands, tmp_path / "out" / "run.txt")
    # the reference way: one padded row per (qid, docid), scored in one batch
    rows = [(q, d) for q in cands for d in cands[q]]
    batch = {"query": torch.tensor([_padlist(queries[q], Q) for q, _ in rows], device=DEV),
             "posdoc": torch.tensor([_padlist(docs[d], D) for _, d in rows], device=DEV),
             "query_idf": torch.tensor([_padlist(idf[q], Q, 0.0) for q, _ in rows], dtype=torch.float32, device=DEV)}
    with torch.no_grad():
        scores = rr.test(batch).view(-1).cpu().numpy()
    want = {}
    for (q, d), s in zip(rows, scores):
        want.setdefault(q, {})[d] = s.astype(np.float16).item()
    assert preds == want
    write_trec_run(want, tmp_path / "ref.txt")
    assert (tmp_path / "out" / "run.txt").read_text() == (tmp_path / "ref.txt").read_text()
    with pytest.raises(KeyError, match="none features"):
        RunPredictor(asm).predict(rr, {"100": ["nosuchdoc"]})


def _bert_rows(query, doc, P, L, maxqlen, passagelen, stride, padq, CLS=101, SEP=102, PAD=0):
    """bertpassage.py:203-232 (sliding windows) + 268-284 (_prepare_bert_input) on token-id lists."""
    passages = [doc[i:i + passagelen] for i in range(0, len(doc), stride)]
    passages = passages[:P] if len(passages) > P else passages + [[PAD] for _ in range(P - len(passages))]
    q = query[:maxqlen] if len(query) > maxqlen else (_padlist(query, maxqlen, PAD) if padq else list(query))
    rows = []
    for psg in passages:
        psg = psg[: L - len(q) - 3]
        line = [CLS] + q + [SEP] + list(psg) + [SEP]
        padded = _padlist(line, L, PAD)
        mask = [1 if t != PAD else 0 for t in line] + [0] * (len(padded) - len(line))
        seg = [0] * (len(q) + 2) + [1] * (len(padded) - len(q) - 2)
        rows.append((padded, mask[:L], seg[:L]))
    return rows


@pytest.mark.parametrize("padq", [False, True])
def test_assemble_bert_pairs_equals_the_reference_extractor_logic(padq):
    from capreolus_b200.predict import BertPairAssembler, PackedIdStore

    P, L, maxqlen, passagelen, stride, V = 4, 48, 8, 20, 15, 30522
    rng = np.random.default_rng(12)
    queries = {str(i): rng.integers(1000, V, size=rng.integers(1, 14)).tolist() for i in range(6)}
    docs = {f"d{i}": rng.integers(1000, V, size=n).tolist() for i, n in enumerate([0, 1, 14, 15, 16, 40, 61, 200, 33, 75])}
    asm = BertPairAssembler(PackedIdStore.from_lists(queries), PackedIdStore.from_lists(docs), maxqlen, L, P, passagelen, stride, DEV, padq=padq)
    qi = rng.integers(0, len(queries), size=64).astype(np.int32)
    di = rng.integers(0, len(docs), size=64).astype(np.int32)
    out = asm.assemble(torch.from_numpy(qi).to(DEV), torch.from_numpy(di).to(DEV))
    qn, dn = list(queries), list(docs)
    for n in range(64):
        want = _bert_rows(queries[qn[qi[n]]], docs[dn[di[n]]], P, L, maxqlen, passagelen, stride, padq)
        for p, (ids, mask, seg) in enumerate(want):
            assert out["pos_bert_input"][n, p].tolist() == ids, (n, p)
            assert out["pos_mask"][n, p].tolist() == mask, (n, p)
            assert out["pos_seg"][n, p].tolist() == seg, (n, p)


def test_run_predictor_drives_a_bert_reranker_from_a_packed_store():
    """RunPredictor + BertPairAssembler + PTBERTMaxP == scoring the rows built by the reference extractor logic."""
    from capreolus_b200 import reranker as R
    from capreolus_b200.predict import BertPairAssembler, PackedIdStore, RunPredictor

    P, L, maxqlen, passagelen, stride = 2, 48, 6, 24, 20
    cfg = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=1000, max_position_embeddings=64)
    rng = np.random.default_rng(3)
    queries = {str(i): rng.integers(200, 1000, size=rng.integers(1, 9)).tolist() for i in range(4)}
    docs = {f"d{i}": rng.integers(200, 1000, size=rng.integers(1, 70)).tolist() for i in range(30)}

    class Ext:
        embeddings = None
        config = {"numpassages": P, "maxseqlen": L}

    torch.manual_seed(0)
    rr = R.PTBERTMaxP(dict(pretrained=cfg), provide={"extractor": Ext()})
    rr.build_model().to(DEV).eval()
    dn = list(docs)
    cands = {q: [dn[j] for j in rng.choice(len(dn), size=7, replace=False)] for q in queries}
    asm = BertPairAssembler(PackedIdStore.from_lists(queries), PackedIdStore.from_lists(docs), maxqlen, L, P, passagelen, stride, DEV)
    preds = RunPredictor(asm, chunk=10).predict(rr, cands)
    rows = [(q, d) for q in cands for d in cands[q]]
    built = [_bert_rows(queries[q], docs[d], P, L, maxqlen, passagelen, stride, False) for q, d in rows]
    batch = {k: torch.tensor([[r[i] for r in b] for b in built], device=DEV) for i, k in enumerate(("pos_bert_input", "pos_mask", "pos_seg"))}
    with torch.no_grad():
        scores = rr.test(batch).view(-1).cpu().numpy()
    for (q, d), s in zip(rows, scores):
        assert preds[q][d] == s.astype(np.float16).item()

"""GPU: engine 3 of the KNRM kernel (csrc/simtc3.cuh, csrc/knrm_tc3.cu) -- the term-frequency pre-pass bit-exactly against numpy,
and the scoring kernel (term-frequency documents, cosines pooled straight from tensor memory) against the CPU oracle, against the
same kernel with the pre-pass switched to the identity (CAPR_KNRM_TF=0) and against the round-1 tensor-core engine."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


class Extractor:
    def __init__(self, table, Q, D):
        self.embeddings = table
        self.config = {"maxqlen": Q, "maxdoclen": D}


def _numpy_dedup(doc):
    """first-occurrence-ordered distinct tokens + counts of every row (the specification of capr_tf_dedup)."""
    B, D = doc.shape
    ids, cnt, nd = np.zeros((B, D), np.int32), np.zeros((B, D), np.uint16), np.zeros(B, np.int32)
    for b in range(B):
        seen = {}
        for t in doc[b].tolist():
            t = int(np.clip(t, -2147483647, 2147483647))
            seen[t] = seen.get(t, 0) + 1
        keys = list(seen)  # dicts keep insertion order = first occurrence
        nd[b] = len(keys)
        ids[b, : len(keys)] = keys
        cnt[b, : len(keys)] = [seen[k] for k in keys]
    return ids, cnt, nd


@pytest.mark.parametrize("B,D,V", [(1, 1, 10), (7, 77, 50), (64, 512, 30000), (33, 1024, 300), (5, 800, 2), (300, 128, 1000)])
def test_tf_dedup_matches_numpy_bit_for_bit(B, D, V):
    from capreolus_b200 import _lib, synthetic

    rng = np.random.default_rng(B * 1000 + D)
    doc = synthetic.zipf_ids(rng, (B, D), max(V, 2))
    doc[rng.random((B, D)) < 0.15] = 0  # pads anywhere (not only trailing)
    doc[rng.random((B, D)) < 0.05] = -rng.integers(1, 9)  # OOV ids
    if B > 2:
        doc[1, :] = 0  # an all-pad document
        doc[2, :] = 7  # one token repeated D times
    if B > 3:
        doc[3, 0] = 2 ** 40  # beyond int32: clamped
    d = torch.from_numpy(doc).to(DEV)
    ids = torch.full((B, D), -99, dtype=torch.int32, device=DEV)
    cnt = torch.full((B, D), 77, dtype=torch.int16, device=DEV)
    nd = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    _lib.check(_lib.lib().capr_tf_dedup(d.data_ptr(), B, D, ids.data_ptr(), cnt.data_ptr(), nd.data_ptr(), None))
    want_ids, want_cnt, want_nd = _numpy_dedup(doc)
    assert np.array_equal(nd.cpu().numpy(), want_nd)
    assert np.array_equal(ids.cpu().numpy(), want_ids)
    assert np.array_equal(cnt.cpu().numpy().view(np.uint16), want_cnt)
    assert np.all(want_cnt.astype(np.int64).sum(axis=1) == D)


def _knrm(table, Q, D, cfg=None, seed=0):
    from capreolus_b200 import reranker as R

    rr = R.KNRM(cfg or {}, provide={"extractor": Extractor(table, Q, D)})
    torch.manual_seed(seed)
    model = rr.build_model().eval()
    with torch.no_grad():
        model.combine[0].weight.mul_(0.05)
    return rr, model


SHAPES = [  # B, Q, D, V, E
    (16, 32, 512, 3000, 300),   # the benchmark shape
    (5, 32, 128, 500, 300),     # exactly one unit
    (9, 7, 129, 200, 64),       # one token into the second unit; a single K atom
    (4, 32, 1024, 5000, 300),   # eight units: the accumulator ring wraps twice per pair
    (6, 20, 800, 3000, 300),    # the reference extractor's default maxdoclen
    (3, 5, 77, 211, 36),        # nothing a multiple of a tile size
    (1, 1, 1, 10, 16),          # degenerate
    (311, 32, 300, 40, 48),     # tiny vocabulary: heavy duplication (few distinct tokens), more pairs than SMs
]


@pytest.mark.parametrize("B,Q,D,V,E", SHAPES)
def test_engine3_matches_oracle_identity_prepass_and_engine2(B, Q, D, V, E, monkeypatch):
    import importlib

    from capreolus_b200 import synthetic
    from oracle import restated

    common = importlib.import_module("capreolus_b200.reranker.common")
    table = synthetic.embedding_table(V, E, seed=3)
    batch = synthetic.parity_batch(B, Q, D, V, seed=B + D, oov=True)
    cpu = {k: torch.from_numpy(v) for k, v in batch.items()}
    gpu = {k: v.to(DEV) for k, v in cpu.items()}
    rr, model = _knrm(table, Q, D)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.to(DEV)
    with torch.no_grad():
        want = restated.knrm_forward(state, torch.from_numpy(table), cpu["posdoc"], cpu["query"]).view(-1).numpy()
        monkeypatch.setattr(common, "ENGINE", "tc3")
        got = rr.test(gpu).cpu().numpy()
        again = rr.test(gpu).cpu().numpy()
        monkeypatch.setenv("CAPR_KNRM_TF", "0")  # identity pre-pass: every position its own token
        ident = rr.test(gpu).cpu().numpy()
        monkeypatch.delenv("CAPR_KNRM_TF")
        monkeypatch.setenv("CAPR_SIM3_QBUFS", "2")  # the other shared-memory layout: two query buffers, a shorter ring
        one_q = rr.test(gpu).cpu().numpy()
        monkeypatch.delenv("CAPR_SIM3_QBUFS")
        monkeypatch.setattr(common, "ENGINE", "tc")
        e2 = rr.test(gpu).cpu().numpy()
    assert rel_err(got, want) < TOL
    assert np.array_equal(got, again)  # bit-reproducible
    assert np.array_equal(got, one_q)  # the layout does not change the arithmetic
    assert rel_err(ident, want) < TOL
    assert rel_err(got, e2, floor=1e-3) < 1e-4  # both engines sit ~1e-6 from the reference


def test_engine3_variants_and_features(monkeypatch):
    """singlefc=False / scoretanh, K = 16 kernels (the KT = 16 instantiation) and the feature output."""
    import importlib

    from capreolus_b200 import synthetic
    from oracle import restated

    monkeypatch.setattr(importlib.import_module("capreolus_b200.reranker.common"), "ENGINE", "tc3")
    B, Q, D, V, E = 12, 32, 512, 3000, 300
    table = synthetic.embedding_table(V, E, seed=3)
    batch = synthetic.parity_batch(B, Q, D, V, seed=11, oov=True)
    cpu = {k: torch.from_numpy(v) for k, v in batch.items()}
    gpu = {k: v.to(DEV) for k, v in cpu.items()}
    for cfg in ({"singlefc": False}, {"scoretanh": True}):
        rr, model = _knrm(table, Q, D, cfg)
        state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model.to(DEV)
        with torch.no_grad():
            want = restated.knrm_forward(state, torch.from_numpy(table), cpu["posdoc"], cpu["query"], singlefc=cfg.get("singlefc", True),
                                         scoretanh=cfg.get("scoretanh", False)).view(-1).numpy()
            got = rr.test(gpu).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-4)


def test_engine3_shards_and_chunked_workspace_are_bitwise_identical(monkeypatch):
    """A pair's score does not depend on which launch / chunk / CTA it lands in: slices of the batch, and a workspace that only
    holds a few pairs at a time (the call loops), give the bits of the single call."""
    import importlib

    from capreolus_b200 import _lib, synthetic

    monkeypatch.setattr(importlib.import_module("capreolus_b200.reranker.common"), "ENGINE", "tc3")
    B, Q, D, V, E = 700, 32, 512, 3000, 300
    table = synthetic.embedding_table(V, E, seed=0)
    gpu = {k: torch.from_numpy(v).to(DEV) for k, v in synthetic.throughput_batch(B, Q, D, V, seed=5).items()}
    gpu["posdoc"][5, 100:] = 0  # ragged documents: fewer distinct tokens than a unit holds / a single token / nothing
    gpu["posdoc"][6, 1:] = 0
    gpu["posdoc"][7, :] = 0
    rr, model = _knrm(table, Q, D)
    model.to(DEV)
    with torch.no_grad():
        whole = rr.test(gpu)
        parts = torch.cat([rr.test({k: v[lo:lo + 233] for k, v in gpu.items()}) for lo in range(0, B, 233)])
        assert torch.equal(whole, parts)
        # a small workspace: 50 pairs at a time
        lib = _lib.lib()
        small = torch.empty(lib.capr_tf_workspace_bytes(50, D), dtype=torch.uint8, device=DEV)
        hi, lo = model._prepared.get_bf16()
        mu, sigma = model.kernels.stacked()
        fc1 = model.combine[0]
        out = torch.empty(B, dtype=torch.float32, device=DEV)
        _lib.check(lib.capr_knrm_forward_tf(gpu["query"].data_ptr(), gpu["posdoc"].data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], E,
                                            hi.shape[1], mu.data_ptr(), sigma.data_ptr(), mu.shape[0], fc1.weight.data_ptr(), fc1.bias.data_ptr(), 0, None,
                                            None, 0, out.data_ptr(), None, small.data_ptr(), small.numel(), torch.cuda.current_stream().cuda_stream))
        assert torch.equal(out, whole.view(-1))
        tiny = torch.empty(256, dtype=torch.uint8, device=DEV)
        rc = lib.capr_knrm_forward_tf(gpu["query"].data_ptr(), gpu["posdoc"].data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], E, hi.shape[1],
                                      mu.data_ptr(), sigma.data_ptr(), mu.shape[0], fc1.weight.data_ptr(), fc1.bias.data_ptr(), 0, None, None, 0,
                                      out.data_ptr(), None, tiny.data_ptr(), tiny.numel(), None)
        assert rc == _lib.BAD_SHAPE

"""GPU parity: the CUDA path (through the C ABI) against (a) the goldens the unmodified reference produced and
(b) the pinned CPU oracle on fresh seeded inputs.  Tolerance: 1e-3 relative (BASELINE.json north_star), fp32."""
import numpy as np
import pytest
import torch

from conftest import golden_state, golden_table, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3
SHAPES = ["full", "small", "odd"]
DEV = "cuda:0"


class Extractor:
    def __init__(self, table, Q, D):
        self.embeddings = table
        self.config = {"maxqlen": Q, "maxdoclen": D}


def _batch(g, dev=DEV):
    return {k: torch.from_numpy(g[k]).to(dev) for k in ("query", "posdoc", "negdoc", "query_idf")}


def _build(cls, g, variant, cfg):
    from capreolus_b200 import reranker as R

    B, Q, D, V, E = (int(x) for x in g["shape"])
    rr = getattr(R, cls)(cfg, provide={"extractor": Extractor(golden_table(g), Q, D)})
    model = rr.build_model()
    missing, unexpected = model.load_state_dict(golden_state(g, variant), strict=False)
    assert not unexpected, unexpected  # every reference key exists in the drop-in module
    assert all("embedding" in k for k in missing), missing
    model.to(DEV).eval()
    return rr, model


KNRM_CFG = {
    "default": dict(gradkernels=True, scoretanh=False, singlefc=True, finetune=False),
    "twofc": dict(gradkernels=True, scoretanh=False, singlefc=False, finetune=False),
    "tanh": dict(gradkernels=True, scoretanh=True, singlefc=True, finetune=False),
}
DRMM_CFG = {
    "default": dict(nbins=29, nodes=5, histType="LCH", gateType="IDF"),
    "nh": dict(nbins=29, nodes=5, histType="NH", gateType="IDF"),
    "ch_tv": dict(nbins=11, nodes=7, histType="CH", gateType="TV"),
}
PACRR_CFG = {
    "default": dict(mingram=1, maxgram=3, nfilters=32, idf=True, kmax=2, combine=32, nonlinearity="relu"),
    "noidf_tanh": dict(mingram=1, maxgram=3, nfilters=32, idf=False, kmax=2, combine=32, nonlinearity="tanh"),
    "wide": dict(mingram=2, maxgram=3, nfilters=16, idf=True, kmax=3, combine=24, nonlinearity="none"),
}


@pytest.mark.parametrize("shape", SHAPES)
def test_simmat_matches_reference(shape):
    g = load_golden(f"knrm_{shape}")
    rr, model = _build("KNRM", g, "default", KNRM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        sim = model.simmat(b["query"][:2], b["posdoc"][:2]).cpu().numpy()
    np.testing.assert_allclose(sim, g["sim_first2"], atol=3e-6)
    # exact structural zeros: padded rows / columns
    q, d = g["query"][:2], g["posdoc"][:2]
    assert np.all(sim[(q == 0)[:, :, None] & np.ones_like(sim, bool)] == 0)
    assert np.all(sim[np.ones_like(sim, bool) & (d == 0)[:, None, :]] == 0)


@pytest.fixture(params=["tc", "tc3", "ffma"])
def engine(request, monkeypatch):
    """Run a test once per cosine-tile engine (tcgen05 tensor cores: "tc" = engine 2, the default; "tc3" = KNRM on engine 3 with
    term-frequency documents pooled from tensor memory, the other models unchanged / fp32 CUDA cores)."""
    import importlib

    monkeypatch.setattr(importlib.import_module("capreolus_b200.reranker.common"), "ENGINE", request.param)
    return request.param


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", list(KNRM_CFG))
def test_knrm_scores_match_reference(shape, variant, engine):
    g = load_golden(f"knrm_{shape}")
    rr, model = _build("KNRM", g, variant, KNRM_CFG[variant])
    b = _batch(g)
    with torch.no_grad():
        pos, neg = rr.score(b)
        assert torch.equal(rr.test(b), pos)
    assert pos.shape == (g["query"].shape[0],)
    if variant == "tanh":  # saturated at +-1: compare absolutely
        np.testing.assert_allclose(pos.cpu().numpy(), g["tanh/pos"], atol=1e-4)
        return
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos"]) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg"]) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_knrm_features_match_reference_soft_tf(shape):
    g = load_golden(f"knrm_{shape}")
    rr, model = _build("KNRM", g, "default", KNRM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        feats = model.kernel_features(b["posdoc"], b["query"]).cpu().numpy()
    live = g["row_live"][:, None, :]  # [B,1,Q]
    want = (np.where(live, np.log(g["soft_tf"].astype(np.float64) + 1e-6), 0.0)).sum(axis=2)  # [B,K]
    assert rel_err(feats, want) < TOL


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", list(DRMM_CFG))
def test_drmm_scores_match_reference(shape, variant, engine):
    """Inputs with exact matches are checked against the reference evaluated with float64 cosines (`pos64`): the fp32
    reference's own bin for identical tokens is rounding noise (tests/test_oracle.py documents it; DESIGN.md 'Exact
    matches').  Inputs without shared terms are checked against the plain fp32 reference."""
    g = load_golden(f"drmm_{shape}")
    rr, model = _build("DRMM", g, variant, DRMM_CFG[variant])
    b = _batch(g)
    with torch.no_grad():
        pos, neg = rr.score(b)
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos64"]) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg64"]) < TOL
    dj = {k: torch.from_numpy(g[f"disjoint/{k}"].astype(np.int64) if g[f"disjoint/{k}"].dtype == np.int32 else g[f"disjoint/{k}"]).to(DEV)
          for k in ("query", "posdoc", "negdoc", "query_idf")}
    with torch.no_grad():
        pos, neg = rr.score(dj)
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/disjoint_pos"]) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/disjoint_neg"]) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_drmm_histogram_matches_reference(shape, engine):
    g = load_golden(f"drmm_{shape}")
    rr, model = _build("DRMM", g, "default", DRMM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        hist = model._hist_map(b["query"], b["posdoc"]).cpu().numpy()
    assert hist.shape == g["hist64"].shape
    # bin counts are integers: a flip happens only when a cosine sits within fp32 rounding of a bin edge
    counts_got, counts_want = np.rint(np.exp(hist)), np.rint(np.exp(g["hist64"]))
    flips = np.abs(counts_got - counts_want).sum() / 2
    assert flips <= max(2, 1e-5 * counts_want.sum()), flips
    assert np.array_equal(counts_got[:, :, -1], counts_want[:, :, -1])  # the exact-match slot (DRMM.py:66)
    # against the fp32 reference: identical except for its coin-flip placement of exact matches in the last regular bin
    c32 = np.rint(np.exp(g["hist"]))
    assert np.abs(counts_got[:, :, :28] - c32[:, :, :28]).sum() / 2 <= max(2, 1e-5 * c32.sum())
    assert np.all(np.abs(counts_got[:, :, 28] - c32[:, :, 28]) <= counts_got[:, :, 29] - 1 + 1e-6)


@pytest.mark.parametrize("shape", ["full", "small"])
def test_drmm_scores_vs_fp32_reference_lie_in_the_exact_match_envelope(shape):
    """Score-level statement of how far the CUDA path sits from the PLAIN fp32 reference (`default/pos`, not `pos64`) on the zipf
    PARITY set.  The only difference is where each exact match (identical in-vocabulary query / doc token) lands: the reference's
    fp32 self-cosine is 1 +- 1 ulp by rounding noise, so `s < 1.0` (DRMM.py:63-65) puts it in or out of bin 28 by coin flip; the
    kernel always counts it in.  For every pair the fp32 reference score must therefore lie in the interval spanned by moving, per
    query row, 0..n_exact of the kernel's bin-28 counts out of that bin (gates are >= 0 and do not depend on the histogram).  Also
    reports which fraction of the pairs agrees with the fp32 reference within 1e-3 outright."""
    import json
    import os

    g = load_golden(f"drmm_{shape}")
    rr, model = _build("DRMM", g, "default", DRMM_CFG["default"])
    b = _batch(g)
    st = {k: v.detach().cpu().double() for k, v in model.state_dict().items() if "embedding" not in k}
    frac = {}
    for side, doc_key in (("pos", "posdoc"), ("neg", "negdoc")):
        with torch.no_grad():
            got = model(b[doc_key], b["query"], b["query_idf"]).view(-1).cpu().numpy().astype(np.float64)
            hist = model._hist_map(b["query"], b[doc_key]).cpu().double()  # [B,Q,30] = log(count + 1)
        counts = torch.round(torch.exp(hist)) - 1.0
        n_exact = counts[:, :, 29].clone()  # the exact-match slot: identical tokens (the disjoint rows have none)
        Bn, Qn, _ = counts.shape
        query = b["query"].cpu()
        q_mask = (query != 0).double()
        logits = b["query_idf"].cpu().double() * st["gates.weight"].view(()) + (1 - q_mask) * -1e7
        gate = torch.softmax(logits, dim=1)  # [B,Q]

        def z_of(cnt):  # ffw on log(count + 1)   (DRMM.py:71-76,106)
            h = torch.log(cnt + 1.0)
            z = torch.tanh(h @ st["ffw.0.weight"].T + st["ffw.0.bias"])
            return torch.tanh(z @ st["ffw.2.weight"].T + st["ffw.2.bias"]).reshape(Bn, Qn)

        z_lo = z_hi = z_of(counts)
        for k in range(1, int(n_exact.max()) + 1):  # k exact matches of the row fall out of bin 28
            moved = counts.clone()
            take = torch.clamp(torch.full_like(n_exact, float(k)), max=n_exact)
            moved[:, :, 28] -= take
            zk = z_of(moved)
            z_lo, z_hi = torch.minimum(z_lo, zk), torch.maximum(z_hi, zk)
        w, bo = float(st["output_layer.weight"].view(())), float(st["output_layer.bias"].view(()))
        a = (gate * z_lo).sum(1).numpy() * w + bo
        c = (gate * z_hi).sum(1).numpy() * w + bo
        lo, hi = np.minimum(a, c), np.maximum(a, c)
        ref32 = g[f"default/{side}"].astype(np.float64)
        slack = 1e-3 * np.maximum(np.abs(ref32), 1e-3)
        inside = (ref32 >= lo - slack) & (ref32 <= hi + slack)
        assert inside.all(), (side, np.nonzero(~inside)[0][:8], ref32[~inside][:8], lo[~inside][:8], hi[~inside][:8])
        assert np.all((got >= lo - slack) & (got <= hi + slack))  # the kernel's own score is the k = 0 corner
        err = np.abs(got - ref32) / np.maximum(np.abs(ref32), 1e-3)
        frac[side] = {"pairs": int(err.size), "within_1e-3_of_fp32_reference": float((err < 1e-3).mean()), "max_rel_err_vs_fp32_reference": float(err.max()),
                      "pairs_with_exact_matches": int((n_exact.sum(1) > 0).sum()), "max_envelope_width_rel": float(((hi - lo) / np.maximum(np.abs(ref32), 1e-3)).max())}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/drmm_fp32_envelope_{shape}.json", "w") as f:
        json.dump({"golden": f"tests/golden/drmm_{shape}.npz", "stats": frac}, f, indent=1)
    print(shape, json.dumps(frac))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", list(PACRR_CFG))
def test_pacrr_scores_match_reference(shape, variant, engine):
    g = load_golden(f"pacrr_{shape}")
    rr, model = _build("PACRR", g, variant, PACRR_CFG[variant])
    b = _batch(g)
    with torch.no_grad():
        pos, neg = rr.score(b)
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos"]) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg"]) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_pacrr_topk_matches_reference(shape, engine):
    g = load_golden(f"pacrr_{shape}")
    rr, model = _build("PACRR", g, "default", PACRR_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        topk = model.ngram_topk(b["posdoc"], b["query"]).cpu().numpy()
    np.testing.assert_allclose(topk, g["topk"], rtol=1e-4, atol=2e-6)


# ---- fresh inputs against the pinned oracle ---------------------------------------------------------------
def _fresh(cls, oracle_fn, cfg, B, Q, D, V, E, seed, oov=True, **okw):
    """okw are passed to the oracle (e.g. exact_cosines=True for DRMM, see test_drmm_scores_match_reference)."""
    from capreolus_b200 import reranker as R, synthetic
    from oracle import restated

    table = synthetic.embedding_table(V, E, seed=seed)
    batch = synthetic.parity_batch(B, Q, D, V, seed=seed + 1, oov=oov)
    torch.manual_seed(seed)
    rr = getattr(R, cls)(cfg, provide={"extractor": Extractor(table, Q, D)})
    model = rr.build_model().eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "embedding" not in n and "kernels" not in n:
                p.mul_(3.0)  # spread the default init so the scores are not dominated by biases
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cpu = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        want = getattr(restated, oracle_fn)(state, torch.from_numpy(table), cpu["posdoc"], cpu["query"], cpu["query_idf"], **okw).view(-1).numpy()
    model.to(DEV)
    with torch.no_grad():
        got = rr.test({k: v.to(DEV) for k, v in cpu.items()}).cpu().numpy()
    return got, want


@pytest.mark.parametrize("B,Q,D,V,E", [(1, 32, 512, 3000, 300), (2, 4, 800, 1000, 300), (5, 17, 1100, 400, 100), (150, 32, 64, 5000, 300), (3, 1, 1, 50, 16),
                                       (150, 32, 800, 3000, 300), (3, 20, 1024, 600, 50), (4, 32, 513, 800, 128)])
def test_knrm_fresh_shapes(B, Q, D, V, E, engine):
    got, want = _fresh("KNRM", "knrm_forward", KNRM_CFG["default"], B, Q, D, V, E, seed=31)
    assert rel_err(got, want) < TOL


@pytest.mark.parametrize("B,Q,D,V,E", [(1, 32, 512, 3000, 300), (2, 4, 800, 1000, 300), (5, 17, 1100, 400, 100), (150, 32, 64, 5000, 300),
                                       (150, 32, 800, 3000, 300), (3, 20, 1024, 600, 50)])
def test_drmm_fresh_shapes(B, Q, D, V, E, engine):
    got, want = _fresh("DRMM", "drmm_forward", DRMM_CFG["default"], B, Q, D, V, E, seed=41, oov=False, exact_cosines=True)
    assert rel_err(got, want) < TOL


@pytest.mark.parametrize("B,Q,D,V,E", [(1, 32, 512, 3000, 300), (2, 4, 300, 1000, 300), (5, 17, 77, 400, 100), (150, 32, 64, 5000, 300)])
def test_pacrr_fresh_shapes(B, Q, D, V, E, engine):
    got, want = _fresh("PACRR", "pacrr_forward", PACRR_CFG["default"], B, Q, D, V, E, seed=51)
    assert rel_err(got, want) < TOL


@pytest.mark.parametrize("B,Q,D,V,E", [(3, 32, 800, 3000, 300), (2, 20, 1023, 1000, 100), (2, 8, 513, 500, 64), (2, 32, 1530, 2000, 300)])
def test_pacrr_long_documents_are_tiled(B, Q, D, V, E):
    """maxdoclen > 512 (the reference EmbedText extractor's default is 800): capr_pacrr_forward tiles the document, carrying the
    per-row top-k lists from tile to tile -- scores and the top-k features themselves against the oracle."""
    from capreolus_b200 import reranker as R, synthetic
    from oracle import restated

    got, want = _fresh("PACRR", "pacrr_forward", PACRR_CFG["default"], B, Q, D, V, E, seed=61)
    assert rel_err(got, want) < TOL
    table = synthetic.embedding_table(V, E, seed=61)
    batch = {k: torch.from_numpy(v) for k, v in synthetic.parity_batch(B, Q, D, V, seed=62, oov=True).items()}
    torch.manual_seed(61)
    rr = R.PACRR(PACRR_CFG["default"], provide={"extractor": Extractor(table, Q, D)})
    model = rr.build_model().eval()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sim = restated.similarity_matrix(torch.from_numpy(table), batch["query"], batch["posdoc"])
    want_topk = torch.cat([restated.pacrr_ngram_topk(sim, state[f"ngrams.{i}.conv.weight"], state[f"ngrams.{i}.conv.bias"], n, 2)
                           for i, n in enumerate(range(1, 4))], dim=2).numpy()
    model.to(DEV)
    with torch.no_grad():
        got_topk = model.ngram_topk(batch["posdoc"].to(DEV), batch["query"].to(DEV)).cpu().numpy()
    np.testing.assert_allclose(got_topk, want_topk, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("cls,fn,cfg", [("ConvKNRM", "convknrm_forward", {}), ("KNRM", "knrm_forward", {}), ("DRMMTKS", "drmmtks_forward", {})])
def test_default_extractor_doclen_800(cls, fn, cfg):
    """The reference extractor's default maxdoclen = 800 runs on the tensor-core engines of the other models too."""
    got, want = _fresh(cls, fn, cfg, 4, 32, 800, 3000, 300, seed=71, oov=cls != "ConvKNRM")
    assert rel_err(got, want) < TOL


def test_errors_and_edge_cases():
    from capreolus_b200 import reranker as R, synthetic

    table = synthetic.embedding_table(100, 32, seed=0)
    rr = R.KNRM(provide={"extractor": Extractor(table, 40, 16)})
    model = rr.build_model().to(DEV).eval()
    q = torch.zeros(2, 40, dtype=torch.long, device=DEV)
    d = torch.zeros(2, 16, dtype=torch.long, device=DEV)
    with torch.no_grad():
        with pytest.raises(ValueError, match="maxqlen"):
            model(d, q, None)
        # empty batch -> empty result
        assert model(d[:0], q[:0, :8], None).shape == (0, 1)
        # all-pad pair: every query row masked -> score = bias (KNRM.py:51-54)
        s = model(d, q[:, :8], None)
        assert torch.allclose(s.view(-1), model.combine[0].bias.expand(2))
        # non-contiguous / int32 ids are accepted
        q2 = torch.randint(1, 100, (2, 16), device=DEV)[:, ::2]
        d2 = torch.randint(1, 100, (2, 16), device=DEV, dtype=torch.int32)
        a = model(d2, q2, None)
        b = model(d2.long(), q2.contiguous(), None)
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(d.cpu(), q[:, :8].cpu(), None)


def test_sharded_scores_are_bitwise_identical_to_single_call(engine):
    """Per-pair arithmetic does not depend on the batch it is in: shard == whole, bit for bit."""
    from capreolus_b200.sharding import shard_bounds

    g = load_golden("knrm_full")
    rr, model = _build("KNRM", g, "default", KNRM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        whole = rr.test(b)
        for world in (2, 3, 8):
            parts = []
            for rank in range(world):
                lo, hi = shard_bounds(whole.shape[0], rank, world)
                parts.append(rr.test({k: v[lo:hi] for k, v in b.items()}))
            assert torch.equal(torch.cat(parts), whole)


def test_full_size_properties(engine):
    """BASELINE.json configs[1] size (100k pairs, |q|=32, |d|=512): size-independent properties + oracle on a sample."""
    from capreolus_b200 import reranker as R, synthetic
    from oracle import restated

    N, Q, D, V, E = 100_000, 32, 512, 30000, 300
    table = synthetic.embedding_table(V, E, seed=0)
    data = synthetic.throughput_batch(N, Q, D, V, seed=2)
    rr = R.KNRM(provide={"extractor": Extractor(table, Q, D)})
    torch.manual_seed(0)
    model = rr.build_model().eval()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.to(DEV)
    gpu = {k: torch.from_numpy(v).to(DEV) for k, v in data.items()}
    gpu["posdoc"][7] = gpu["posdoc"][3]
    gpu["query"][7] = gpu["query"][3]  # a duplicated pair
    with torch.no_grad():
        s = rr.test(gpu)
        assert s.shape == (N,) and torch.isfinite(s).all()
        assert s[7] == s[3]
        perm = torch.randperm(N, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
        s_perm = rr.test({k: v[perm] for k, v in gpu.items()})
        assert torch.equal(s_perm, s[perm])  # permutation equivariance, bitwise
        idx = torch.arange(0, N, N // 48)[:48]
        want = restated.knrm_forward(state, torch.from_numpy(table), gpu["posdoc"][idx].cpu(), gpu["query"][idx].cpu()).view(-1).numpy()
    assert rel_err(s[idx].cpu().numpy(), want) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_pacrr_conv_on_tensor_cores_engine_matches_reference(shape, monkeypatch):
    """The opt-in engine 3 (CAPR_PACRR_CONV=tc3: im2col + tcgen05 conv) computes the same scores and top-k."""
    monkeypatch.setenv("CAPR_PACRR_CONV", "tc3")
    g = load_golden(f"pacrr_{shape}")
    for variant in ("default", "noidf_tanh"):
        rr, model = _build("PACRR", g, variant, PACRR_CFG[variant])
        b = _batch(g)
        with torch.no_grad():
            pos, neg = rr.score(b)
        assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos"]) < TOL
        assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg"]) < TOL
    rr, model = _build("PACRR", g, "default", PACRR_CFG["default"])
    with torch.no_grad():
        topk = model.ngram_topk(_batch(g)["posdoc"], _batch(g)["query"]).cpu().numpy()
    np.testing.assert_allclose(topk, g["topk"], rtol=1e-4, atol=2e-5)

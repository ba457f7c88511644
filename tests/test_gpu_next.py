"""GPU parity for the SURVEY.md §8(f) rank-1 models, DRMMTKS and ConvKNRM: the CUDA path (through the C ABI) against the
goldens the unmodified reference produced and against the pinned CPU oracle on fresh seeded inputs.  Tolerance 1e-3."""
import numpy as np
import pytest
import torch

from conftest import golden_state, golden_table, load_golden, rel_err
from test_gpu_parity import DEV, SHAPES, TOL, Extractor, _batch, _build, _fresh

pytestmark = pytest.mark.gpu

DRMMTKS_CFG = {
    "default": dict(topk=10, gateType="IDF", freezeemb=True),
    "k3": dict(topk=3, gateType="IDF", freezeemb=True),
    "k20": dict(topk=20, gateType="IDF", freezeemb=False),
}
CONVKNRM_CFG = {
    "default": dict(gradkernels=True, maxngram=3, crossmatch=True, filters=128, scoretanh=False, singlefc=True),
    "nocross_twofc": dict(gradkernels=True, maxngram=2, crossmatch=False, filters=48, scoretanh=False, singlefc=False),
    "uni_tanh": dict(gradkernels=False, maxngram=1, crossmatch=True, filters=128, scoretanh=True, singlefc=True),
}


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", list(DRMMTKS_CFG))
def test_drmmtks_scores_match_reference(shape, variant):
    g = load_golden(f"drmmtks_{shape}")
    rr, model = _build("DRMMTKS", g, variant, DRMMTKS_CFG[variant])
    b = _batch(g)
    with torch.no_grad():
        pos, neg = rr.score(b)
        assert torch.equal(rr.test(b), pos)
    assert pos.shape == (g["query"].shape[0],)
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos"]) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg"]) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_drmmtks_topk_matches_reference(shape):
    g = load_golden(f"drmmtks_{shape}")
    rr, model = _build("DRMMTKS", g, "default", DRMMTKS_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        top = model.topk_similarities(b["posdoc"], b["query"]).cpu().numpy()
    assert top.shape == g["topk"].shape
    assert np.all(np.diff(top, axis=2) <= 0)  # descending, like torch.topk
    np.testing.assert_allclose(top, g["topk"], atol=3e-6)


@pytest.mark.parametrize("B,Q,D,V,E", [(1, 32, 512, 3000, 300), (2, 4, 300, 1000, 300), (5, 17, 77, 400, 100), (150, 32, 64, 5000, 300), (3, 1, 10, 50, 16),
                                       (2, 4, 800, 1000, 300), (3, 32, 1000, 2000, 300), (2, 8, 1024, 500, 50)])
def test_drmmtks_fresh_shapes(B, Q, D, V, E):
    got, want = _fresh("DRMMTKS", "drmmtks_forward", DRMMTKS_CFG["default"], B, Q, D, V, E, seed=61)
    assert rel_err(got, want) < TOL


def test_drmmtks_errors():
    from capreolus_b200 import reranker as R, synthetic

    table = synthetic.embedding_table(100, 32, seed=0)
    q = torch.ones(2, 8, dtype=torch.long, device=DEV)
    d = torch.ones(2, 16, dtype=torch.long, device=DEV)
    idf = torch.zeros(2, 8, device=DEV)
    with torch.no_grad():
        tv = R.DRMMTKS(dict(gateType="TV"), provide={"extractor": Extractor(table, 8, 16)}).build_model().to(DEV).eval()
        with pytest.raises(ValueError, match="gateType"):
            tv(d, q, idf)
        big = R.DRMMTKS(dict(topk=17), provide={"extractor": Extractor(table, 8, 16)}).build_model().to(DEV).eval()
        with pytest.raises(ValueError, match="topk"):  # torch.topk raises in the reference when k > maxdoclen
            big(d, q, idf)
        ok = R.DRMMTKS(provide={"extractor": Extractor(table, 8, 16)}).build_model().to(DEV).eval()
        assert ok(d[:0], q[:0], idf[:0]).shape == (0, 1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", list(CONVKNRM_CFG))
def test_convknrm_scores_match_reference(shape, variant):
    g = load_golden(f"convknrm_{shape}")
    rr, model = _build("ConvKNRM", g, variant, CONVKNRM_CFG[variant])
    b = _batch(g)
    with torch.no_grad():
        pos, neg = rr.score(b)
        assert torch.equal(rr.test(b), pos)
    assert pos.shape == (g["query"].shape[0],)
    if variant == "uni_tanh":  # tanh output: compare absolutely
        np.testing.assert_allclose(pos.cpu().numpy(), g[f"{variant}/pos"], atol=1e-3)
        np.testing.assert_allclose(neg.cpu().numpy(), g[f"{variant}/neg"], atol=1e-3)
        return
    assert rel_err(pos.cpu().numpy(), g[f"{variant}/pos"], floor=1e-2) < TOL
    assert rel_err(neg.cpu().numpy(), g[f"{variant}/neg"], floor=1e-2) < TOL


@pytest.mark.parametrize("shape", SHAPES)
def test_convknrm_features_match_reference(shape):
    g = load_golden(f"convknrm_{shape}")
    rr, model = _build("ConvKNRM", g, "default", CONVKNRM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        feats = model.kernel_features(b["posdoc"], b["query"]).cpu().numpy()
    assert feats.shape == g["feats"].shape
    assert rel_err(feats, g["feats"], floor=1e-1) < TOL


def test_convknrm_chunked_equals_whole(monkeypatch):
    """The rep table is staged chunk by chunk: a 7-pair chunk must give bit-identical scores to one 64-pair chunk."""
    import importlib

    M = importlib.import_module("capreolus_b200.reranker.ConvKNRM")  # (the package attribute of that name is the class)

    g = load_golden("convknrm_full")
    b = _batch(g)
    rr, model = _build("ConvKNRM", g, "default", CONVKNRM_CFG["default"])
    with torch.no_grad():
        whole = rr.test(b)
        monkeypatch.setattr(M, "CHUNK", 7)
        model._ws = None
        chunked = rr.test(b)
    assert torch.equal(whole, chunked)


@pytest.mark.parametrize("B,Q,D,V,E", [(1, 32, 512, 3000, 300), (2, 4, 300, 1000, 300), (5, 17, 77, 400, 100), (150, 32, 64, 5000, 300), (3, 1, 2, 50, 16),
                                       (3, 32, 800, 3000, 300), (2, 20, 1000, 2000, 300)])  # the extractor's default maxdoclen = 800; up to 1024
def test_convknrm_fresh_shapes(B, Q, D, V, E):
    got, want = _fresh("ConvKNRM", "convknrm_forward", CONVKNRM_CFG["default"], B, Q, D, V, E, seed=71, oov=False)
    assert rel_err(got, want, floor=1e-2) < TOL


def test_convknrm_projection_follows_weight_updates():
    """The projected tables are derived data: an in-place change of a conv weight must be picked up by the next call."""
    g = load_golden("convknrm_small")
    rr, model = _build("ConvKNRM", g, "default", CONVKNRM_CFG["default"])
    b = _batch(g)
    with torch.no_grad():
        s0 = rr.test(b).clone()
        model.convs[1][0].weight.mul_(-1.0)
        s1 = rr.test(b)
        assert not torch.equal(s0, s1)
        model.convs[1][0].weight.mul_(-1.0)
        assert torch.equal(rr.test(b), s0)


# ---- CEDR-KNRM -------------------------------------------------------------------------------------
class BertExtractor:
    def __init__(self, P, L, maxqlen):
        self.embeddings = None
        self.config = {"numpassages": P, "maxseqlen": L, "maxqlen": maxqlen}


def _build_cedr(name, variant):
    import json

    from capreolus_b200 import reranker as R

    g = load_golden(f"cedrknrm_{name}")
    cfg = json.loads(str(g["config_json"]))
    vcfg = json.loads(str(g[f"{variant}/config_json"]))
    N, P, L, maxqlen = (int(x) for x in g["shape"])
    keep = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "vocab_size", "max_position_embeddings",
            "type_vocab_size", "initializer_range", "layer_norm_eps", "hidden_act")
    torch.manual_seed(int(g["weight_seed"]))
    rr = R.CEDRKNRM(dict(pretrained={k: cfg[k] for k in keep if k in cfg}, **vcfg), provide={"extractor": BertExtractor(P, L, maxqlen)})
    model = rr.build_model()
    tot = sum(float(v.double().abs().sum()) for v in model.bert.state_dict().values() if v.dtype.is_floating_point)
    np.testing.assert_allclose(tot, g["weight_checksum"][0], rtol=1e-9)  # same random-init encoder as the golden's
    missing, unexpected = model.load_state_dict(golden_state(g, variant), strict=False)
    assert not unexpected and all(k.startswith("bert.") or k in ("one", "zero") for k in missing), (missing, unexpected)
    model.to(DEV).eval()
    batch = {k: torch.from_numpy(g[k].astype(np.int64)).to(DEV) for k in ("pos_bert_input", "pos_mask", "pos_seg")}
    return g, rr, model, batch


@pytest.mark.parametrize("name,variants", [("tiny", ["default", "nocls", "clsonly"]), ("mid", ["default", "nocls"]), ("base", ["default", "nocls"])])
def test_cedrknrm_scores_match_reference(name, variants):
    for variant in variants:
        g, rr, model, b = _build_cedr(name, variant)
        scores = rr.test(b).cpu().numpy()
        assert scores.shape == g[f"{variant}/scores"].shape
        assert rel_err(scores, g[f"{variant}/scores"], floor=1e-2) < TOL, variant


@pytest.mark.parametrize("name", ["tiny", "mid", "base"])
def test_cedrknrm_features_match_reference(name):
    g, rr, model, b = _build_cedr(name, "default")
    feats = model.features(b["pos_bert_input"], b["pos_mask"], b["pos_seg"]).cpu().numpy()
    assert feats.shape == g["feats"].shape
    # absolute: cls features are LayerNorm outputs (|x| ~ 1), knrm features are 0.01 * log sums (|x| ~ 1)
    np.testing.assert_allclose(feats, g["feats"], atol=2e-3)


def test_cedrknrm_doc_chunking_is_invisible():
    g, rr, model, b = _build_cedr("tiny", "default")
    whole = rr.test(b)
    model.max_seqs_per_call, model._engine = 3, None  # one document (3 passages) per encoder call
    assert torch.equal(rr.test(b), whole)
    model.train()
    with pytest.raises(NotImplementedError):
        rr.test(b)


def test_cedrknrm_electra_encoder():
    """CEDR-KNRM's reference default encoder is Electra (CEDRKNRM.py:20-27).  electra-base has no embedding projection, so its
    encoder is the BERT encoder under other names: the engine's hidden states are checked against HF ``ElectraModel`` (the module
    the reference calls) on the CPU, and the scores against the oracle's CEDR-KNRM forward on the same state."""
    import copy

    from capreolus_b200 import reranker as R, synthetic
    from oracle import restated

    P, L, maxqlen, N = 3, 48, 6, 5
    ecfg = dict(model_type="electra", hidden_size=64, embedding_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                vocab_size=1000, max_position_embeddings=64)
    torch.manual_seed(21)
    rr = R.CEDRKNRM(dict(pretrained=ecfg, simmat_layers=[0, 1, 2], combine_hidden=16, cls="avg"), provide={"extractor": BertExtractor(P, L, maxqlen)})
    model = rr.build_model().eval()
    assert model.bert.config.model_type == "electra"
    cpu_model = copy.deepcopy(model.bert).eval()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = {k: torch.from_numpy(v) for k, v in synthetic.cedr_batch(N, P, L, maxqlen, vocab=1000, seed=23).items()}
    flat = lambda t: t.reshape(N * P, L)
    with torch.no_grad():
        hf = cpu_model(flat(batch["pos_bert_input"]), attention_mask=flat(batch["pos_mask"]), token_type_ids=flat(batch["pos_seg"])).hidden_states
        want = restated.cedrknrm_forward(state, batch["pos_bert_input"], batch["pos_mask"], batch["pos_seg"], 4, maxqlen, [0, 1, 2], "avg", 16).view(-1).numpy()
    model.to(DEV)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    hs = model.engine().hidden_states(flat(gb["pos_bert_input"]), flat(gb["pos_mask"]), flat(gb["pos_seg"]), [0, 1, 2]).cpu().numpy()
    real = flat(batch["pos_mask"]).numpy().astype(bool).reshape(-1)  # HF leaves garbage-free but differently-attended pad rows: compare real tokens
    for i in range(3):
        np.testing.assert_allclose(hs[i][real], hf[i].reshape(N * P * L, -1).numpy()[real], atol=2e-3)
    scores = rr.test(gb).cpu().numpy()
    assert rel_err(scores, want, floor=1e-2) < TOL
    with pytest.raises(ValueError):  # an embedding projection is not implemented
        R.CEDRKNRM(dict(pretrained={**ecfg, "embedding_size": 32}), provide={"extractor": BertExtractor(P, L, maxqlen)}).build_model().to(DEV).engine()


# ---- PARADE --------------------------------------------------------------------------------------
def _build_parade(name):
    import json

    from capreolus_b200 import reranker as R

    g = load_golden(f"parade_{name}")
    cfg = json.loads(str(g["config_json"]))
    N, P, L, maxqlen = (int(x) for x in g["shape"])
    keep = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "vocab_size", "max_position_embeddings",
            "type_vocab_size", "initializer_range", "layer_norm_eps", "hidden_act")
    torch.manual_seed(int(g["weight_seed"]))
    rr = R.PTParade(dict(pretrained={k: cfg[k] for k in keep if k in cfg}), provide={"extractor": BertExtractor(P, L, maxqlen)})
    model = rr.build_model()
    tot = sum(float(v.double().abs().sum()) for v in model.bert.state_dict().values() if v.dtype.is_floating_point)
    np.testing.assert_allclose(tot, g["weight_checksum"][0], rtol=1e-9)  # same random-init passage encoder as the golden's
    state = {k[len("state/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state/")}
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not unexpected and all(k.startswith("bert.") for k in missing), (missing, unexpected)
    model.to(DEV).eval()
    batch = {k: torch.from_numpy(g[k].astype(np.int64)).to(DEV) for k in ("pos_bert_input", "pos_mask", "pos_seg")}
    return g, rr, model, batch


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_parade_matches_reference(name):
    g, rr, model, b = _build_parade(name)
    agg = model.aggregate_using_transformer_output(b["pos_bert_input"], b["pos_mask"], b["pos_seg"]).cpu().numpy()
    np.testing.assert_allclose(agg, g["aggregated"], atol=2e-3)  # LayerNorm outputs, |x| ~ 1
    scores = rr.test(b).cpu().numpy()
    assert scores.shape == g["scores"].shape
    assert rel_err(scores, g["scores"], floor=1e-2) < TOL
    # chunking over documents is invisible
    model.max_seqs_per_call, model._engine = int(g["shape"][1]), None
    assert torch.equal(rr.test(b).cpu(), torch.from_numpy(scores))
    model.train()
    with pytest.raises(NotImplementedError):
        rr.test(b)

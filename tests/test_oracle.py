"""CPU: pins oracle/restated.py against the goldens that the UNMODIFIED reference produced
(oracle/make_goldens.py), and -- when /root/reference is present -- against the live reference modules."""
import numpy as np
import pytest
import torch

from conftest import golden_state, golden_table, load_golden, rel_err
from oracle import refshim, restated

SHAPES = ["full", "small", "odd"]
TOL = 2e-5  # restatement vs reference on CPU: same ops, different association only


def _tb(g):
    return {k: torch.from_numpy(g[k]) for k in ("query", "posdoc", "negdoc", "query_idf")}


@pytest.mark.parametrize("shape", SHAPES)
def test_similarity_and_soft_tf_match_reference(shape):
    g = load_golden(f"knrm_{shape}")
    table = torch.from_numpy(golden_table(g))
    tb = _tb(g)
    sim = restated.similarity_matrix(table, tb["query"], tb["posdoc"])
    np.testing.assert_allclose(sim[:2].numpy(), g["sim_first2"], atol=2e-6)
    st = golden_state(g, "default")
    kp = restated.knrm_params_from_state(st)
    soft_tf = restated.rbf_bank(sim, kp["mus"], kp["sigmas"]).sum(dim=3)
    np.testing.assert_allclose(soft_tf.numpy(), g["soft_tf"], rtol=1e-4, atol=1e-5)
    assert np.array_equal((sim.sum(dim=2) != 0).numpy(), g["row_live"])


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant", ["default", "twofc", "tanh"])
def test_knrm_scores(shape, variant):
    g = load_golden(f"knrm_{shape}")
    table = torch.from_numpy(golden_table(g))
    tb = _tb(g)
    st = golden_state(g, variant)
    kw = dict(singlefc=variant != "twofc", scoretanh=variant == "tanh")
    for side, doc in (("pos", "posdoc"), ("neg", "negdoc")):
        got = restated.knrm_forward(st, table, tb[doc], tb["query"], tb["query_idf"], **kw).view(-1).numpy()
        assert rel_err(got, g[f"{variant}/{side}"]) < TOL * 10


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant,kw", [
    ("default", dict(nbins=29, hist_type="LCH", gate_type="IDF")),
    ("nh", dict(nbins=29, hist_type="NH", gate_type="IDF")),
    ("ch_tv", dict(nbins=11, hist_type="CH", gate_type="TV")),
])
def test_drmm_scores(shape, variant, kw):
    g = load_golden(f"drmm_{shape}")
    table = torch.from_numpy(golden_table(g))
    tb = _tb(g)
    st = golden_state(g, variant)
    for side, doc in (("pos", "posdoc"), ("neg", "negdoc")):
        got = restated.drmm_forward(st, table, tb[doc], tb["query"], tb["query_idf"], **kw).view(-1).numpy()
        assert rel_err(got, g[f"{variant}/{side}"]) < 1e-4
    if variant == "default":
        sim = restated.similarity_matrix(table, tb["query"], tb["posdoc"])
        hist = restated.drmm_histogram(sim, tb["posdoc"], 29, "LCH").numpy()
        # counts are discontinuous in the cosine: allow a handful of bin-edge flips
        assert (np.abs(hist - g["hist"]) > 1e-5).mean() < 1e-4


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant,kw", [
    ("default", dict(mingram=1, maxgram=3, kmax=2, idf=True, nonlinearity="relu")),
    ("noidf_tanh", dict(mingram=1, maxgram=3, kmax=2, idf=False, nonlinearity="tanh")),
    ("wide", dict(mingram=2, maxgram=3, kmax=3, idf=True, nonlinearity="none")),
])
def test_pacrr_scores(shape, variant, kw):
    g = load_golden(f"pacrr_{shape}")
    table = torch.from_numpy(golden_table(g))
    tb = _tb(g)
    st = golden_state(g, variant)
    for side, doc in (("pos", "posdoc"), ("neg", "negdoc")):
        got = restated.pacrr_forward(st, table, tb[doc], tb["query"], tb["query_idf"], **kw).view(-1).numpy()
        assert rel_err(got, g[f"{variant}/{side}"]) < 1e-4


def test_losses():
    g = load_golden("losses")
    pos, neg = torch.from_numpy(g["pos"]), torch.from_numpy(g["neg"])
    np.testing.assert_allclose(restated.pair_hinge_loss(pos, neg).numpy(), g["hinge"], rtol=1e-6)
    np.testing.assert_allclose(restated.pair_softmax_loss(pos, neg).numpy(), g["softmax"], rtol=1e-6)


def _hf_bert(g):
    import json

    import transformers

    cfg = transformers.BertConfig(**json.loads(str(g["config_json"])))
    torch.manual_seed(int(g["weight_seed"]))
    model = transformers.BertForSequenceClassification(cfg).eval()
    tot = sum(float(v.double().abs().sum()) for v in model.state_dict().values() if v.dtype.is_floating_point)
    np.testing.assert_allclose(tot, g["weight_checksum"][0], rtol=1e-9)
    return cfg, model


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_bert_restatement_matches_hf_and_reference_maxp(name):
    g = load_golden(f"bert_{name}")
    cfg, model = _hf_bert(g)
    st = {k: v for k, v in model.state_dict().items()}
    ids, mask, seg = (torch.from_numpy(g[k]) for k in ("pos_bert_input", "pos_mask", "pos_seg"))
    N, P, L = ids.shape
    with torch.no_grad():
        logits = restated.bert_logits(st, ids.reshape(N * P, L), mask.reshape(N * P, L), seg.reshape(N * P, L), cfg.num_attention_heads)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=2e-6)
    for agg in ["max", "first", "sum", "avg"]:
        with torch.no_grad():
            got = restated.bert_maxp_forward(st, ids, mask, seg, cfg.num_attention_heads, agg)
        np.testing.assert_allclose(got.numpy(), g[f"{agg}/scores"], rtol=1e-4, atol=2e-6)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_live_reference_agrees_with_goldens():
    """The goldens are reproducible: re-running the unmodified reference KNRM gives the stored scores."""
    ref = refshim.load_rerankers()
    g = load_golden("knrm_small")
    table = golden_table(g)
    B, Q, D, V, E = (int(x) for x in g["shape"])
    rr = ref.KNRM.KNRM(dict(gradkernels=True, scoretanh=False, singlefc=True, finetune=False),
                       provide={"extractor": refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)})
    model = rr.build_model().eval()
    model.load_state_dict(golden_state(g, "default"), strict=False)
    with torch.no_grad():
        pos = rr.test(_tb(g))
    np.testing.assert_allclose(pos.numpy(), g["default/pos"], rtol=1e-5, atol=1e-5)

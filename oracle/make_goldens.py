"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the UNMODIFIED reference
classes (loaded by ``oracle/refshim.py`` from /root/reference, commit 789288c) on seeded synthetic inputs.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_goldens            # all fixtures
    python -m oracle.make_goldens knrm drmm  # a subset

Inputs come from ``capreolus_b200/synthetic.py``; the 36 MB embedding table and BERT-base weights are not
stored -- they are re-derived from their seeds and a checksum of each is stored instead.
"""
from __future__ import annotations

import contextlib
import sys
from pathlib import Path

import numpy as np
import torch

from capreolus_b200 import synthetic
from oracle import refshim

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"
SHAPES = {
    # name: (B, Q, D, V, E, table_seed, input_seed)
    "full": (64, 32, 512, 30000, 300, 0, 1),  # BASELINE.json configs[0]
    "small": (6, 8, 40, 500, 50, 10, 11),
    "odd": (3, 5, 77, 211, 36, 20, 21),  # nothing a multiple of a tile size
}


def table_checksum(table: np.ndarray) -> np.ndarray:
    return np.array([table.astype(np.float64).sum(), np.abs(table.astype(np.float64)).sum(), float(table[-1, -1])])


def _state_np(model, skip=("embedding.weight",)):
    return {f"state/{k}": v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if not any(s in k for s in skip)}


def _inputs(shape_name, oov=True):
    B, Q, D, V, E, tseed, iseed = SHAPES[shape_name]
    table = synthetic.embedding_table(V, E, seed=tseed)
    batch = synthetic.parity_batch(B, Q, D, V, seed=iseed, oov=oov)
    return table, batch


def _common(shape_name, table, batch):
    B, Q, D, V, E, tseed, iseed = SHAPES[shape_name]
    out = {k: (v.astype(np.int32) if v.dtype == np.int64 else v) for k, v in batch.items()}
    out.update(shape=np.array([B, Q, D, V, E]), table_seed=np.array(tseed), input_seed=np.array(iseed),
               table_checksum=table_checksum(table), reference_commit=np.array(refshim.REFERENCE_COMMIT))
    return out


def _t(batch):
    return {k: torch.from_numpy(v) for k, v in batch.items()}


def make_knrm():
    ref = refshim.load_rerankers()
    for shape_name in SHAPES:
        table, batch = _inputs(shape_name)
        B, Q, D, V, E, *_ = SHAPES[shape_name]
        tb = _t(batch)
        out = _common(shape_name, table, batch)
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        for variant, cfg in {
            "default": dict(gradkernels=True, scoretanh=False, singlefc=True, finetune=False),
            "twofc": dict(gradkernels=True, scoretanh=False, singlefc=False, finetune=False),
            "tanh": dict(gradkernels=True, scoretanh=True, singlefc=True, finetune=False),
        }.items():
            torch.manual_seed(100)
            # go through the reference Reranker wrapper: build_model / score / test (KNRM.py:81-101)
            rr = ref.KNRM.KNRM(cfg, provide={"extractor": ext})
            model = rr.build_model().eval()
            if variant == "default" and shape_name != "full":
                # distinct, non-default kernel parameters so mu/sigma plumbing is really checked
                with torch.no_grad():
                    for i, k in enumerate(model.kernels.kernels):
                        k.mu.add_(0.013 * (i - 4))
                        k.sigma.mul_(1.0 + 0.05 * i)
            with torch.no_grad():
                pos, neg = rr.score(tb)
                assert torch.equal(rr.test(tb), pos)
                if variant == "default":
                    sim = model.simmat(tb["query"], tb["posdoc"])
                    kern = model.kernels(sim)
                    soft_tf = kern.sum(dim=3)  # [B,K,Q]
                    out["sim_first2"] = sim[:2].numpy()
                    out["soft_tf"] = soft_tf.numpy()
                    out["row_live"] = (sim.sum(dim=2) != 0.0).numpy()
            out[f"{variant}/pos"] = pos.numpy()
            out[f"{variant}/neg"] = neg.numpy()
            out.update({f"{variant}/{k}": v for k, v in _state_np(model).items()})
        np.savez_compressed(GOLDEN / f"knrm_{shape_name}.npz", **out)
        print("knrm", shape_name, out["default/pos"][:4])


def make_drmm():
    ref = refshim.load_rerankers()
    for shape_name in SHAPES:
        table, batch = _inputs(shape_name, oov=False)  # DRMM.py:109 cannot take negative (OOV) query ids
        B, Q, D, V, E, *_ = SHAPES[shape_name]
        tb = _t(batch)
        dj = synthetic.parity_batch(B, Q, D, V, seed=SHAPES[shape_name][6] + 100, disjoint=True)
        tdj = _t(dj)
        out = _common(shape_name, table, batch)
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        for variant, cfg in {
            "default": dict(nbins=29, nodes=5, histType="LCH", gateType="IDF"),
            "nh": dict(nbins=29, nodes=5, histType="NH", gateType="IDF"),
            "ch_tv": dict(nbins=11, nodes=7, histType="CH", gateType="TV"),
        }.items():
            torch.manual_seed(200)
            rr = ref.DRMM.DRMM(cfg, provide={"extractor": ext})
            model = rr.build_model().eval()
            with torch.no_grad():
                # the default init (|w| <= 0.1 / 0.01) makes every score ~bias; spread the weights so that
                # the histogram and the gate actually matter in the parity check
                model.ffw[0].weight.mul_(10.0)
                model.ffw[2].weight.mul_(10.0)
                model.gates.weight.mul_(40.0)
                pos, neg = rr.score(tb)
                if variant == "default":
                    d_mask = (tb["posdoc"] != 0).float()
                    out["hist"] = model._hist_map(tb["query"], tb["posdoc"], d_mask).numpy()
                # inputs free of in-vocabulary exact matches: the fp32 reference is trustworthy there
                pos_dj, neg_dj = rr.score(tdj)
                # the same reference module, but with its SimilarityMatrix (reference class) holding a float64 copy of
                # the table: cosines in exact arithmetic, everything after the integer counts unchanged (fp32)
                emb64 = torch.nn.Embedding.from_pretrained(torch.from_numpy(table).double())
                fp32_simmat, model.simmat = model.simmat, ref.common.SimilarityMatrix(emb64)
                pos64, neg64 = rr.score(tb)
                if variant == "default":
                    out["hist64"] = model._hist_map(tb["query"], tb["posdoc"], d_mask).numpy()
                model.simmat = fp32_simmat
            out[f"{variant}/pos"] = pos.numpy()
            out[f"{variant}/neg"] = neg.numpy()
            out[f"{variant}/pos64"] = pos64.numpy()
            out[f"{variant}/neg64"] = neg64.numpy()
            out[f"{variant}/disjoint_pos"] = pos_dj.numpy()
            out[f"{variant}/disjoint_neg"] = neg_dj.numpy()
            out.update({f"{variant}/{k}": v for k, v in _state_np(model).items()})
        out.update({f"disjoint/{k}": (v.astype(np.int32) if v.dtype == np.int64 else v) for k, v in dj.items()})
        np.savez_compressed(GOLDEN / f"drmm_{shape_name}.npz", **out)
        print("drmm", shape_name, out["default/pos"][:4])


def make_pacrr():
    ref = refshim.load_rerankers()
    for shape_name in SHAPES:
        table, batch = _inputs(shape_name)
        B, Q, D, V, E, *_ = SHAPES[shape_name]
        tb = _t(batch)
        out = _common(shape_name, table, batch)
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        for variant, cfg in {
            "default": dict(mingram=1, maxgram=3, nfilters=32, idf=True, kmax=2, combine=32, nonlinearity="relu"),
            "noidf_tanh": dict(mingram=1, maxgram=3, nfilters=32, idf=False, kmax=2, combine=32, nonlinearity="tanh"),
            "wide": dict(mingram=2, maxgram=3, nfilters=16, idf=True, kmax=3, combine=24, nonlinearity="none"),
        }.items():
            torch.manual_seed(300)
            rr = ref.PACRR.PACRR(cfg, provide={"extractor": ext})
            model = rr.build_model().eval()
            with torch.no_grad():
                idf_arg = {**tb, "query_idf": refshim.pacrr_idf(tb["query_idf"])}  # PACRR.py:49 work-around
                pos, neg = rr.score(idf_arg)
                if variant == "default":
                    sim = model.simmat(tb["query"], tb["posdoc"])
                    out["topk"] = torch.cat([ng(sim) for ng in model.ngrams], dim=2).numpy()  # [B,Q,6]
            out[f"{variant}/pos"] = pos.numpy()
            out[f"{variant}/neg"] = neg.numpy()
            out.update({f"{variant}/{k}": v for k, v in _state_np(model).items()})
        np.savez_compressed(GOLDEN / f"pacrr_{shape_name}.npz", **out)
        print("pacrr", shape_name, out["default/pos"][:4])


def make_drmmtks():
    """the reference DRMMTKS wrapper (reranker/DRMMTKS.py) on the parity inputs."""
    ref = refshim.load_rerankers()
    for shape_name in SHAPES:
        table, batch = _inputs(shape_name)
        B, Q, D, V, E, *_ = SHAPES[shape_name]
        tb = _t(batch)
        out = _common(shape_name, table, batch)
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        for variant, cfg in {
            "default": dict(topk=10, gateType="IDF", freezeemb=True),
            "k3": dict(topk=3, gateType="IDF", freezeemb=True),
            "k20": dict(topk=20, gateType="IDF", freezeemb=False),
        }.items():
            torch.manual_seed(400)
            rr = ref.DRMMTKS.DRMMTKS(cfg, provide={"extractor": ext})
            model = rr.build_model().eval()
            with torch.no_grad():
                # default init (|w| <= 0.1 / 0.01) makes every score ~bias: spread the weights (as for DRMM above)
                model.ffw[0].weight.mul_(10.0)
                model.gates.weight.mul_(40.0)
                pos, neg = rr.score(tb)
                assert torch.equal(rr.test(tb), pos)
                if variant == "default":
                    out["topk"] = torch.topk(model.simmat(tb["query"], tb["posdoc"]), k=10, dim=-1)[0].numpy()
            out[f"{variant}/pos"] = pos.numpy()
            out[f"{variant}/neg"] = neg.numpy()
            out.update({f"{variant}/{k}": v for k, v in _state_np(model).items()})
        np.savez_compressed(GOLDEN / f"drmmtks_{shape_name}.npz", **out)
        print("drmmtks", shape_name, out["default/pos"][:4])


def make_convknrm():
    """the reference ConvKNRM wrapper (reranker/ConvKNRM.py) on the parity inputs
    (no OOV ids: ConvKNRM.py:44-45 indexes the table with the raw ids)."""
    ref = refshim.load_rerankers()
    for shape_name in SHAPES:
        table, batch = _inputs(shape_name, oov=False)
        B, Q, D, V, E, *_ = SHAPES[shape_name]
        tb = _t(batch)
        out = _common(shape_name, table, batch)
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        for variant, cfg in {
            "default": dict(gradkernels=True, maxngram=3, crossmatch=True, filters=128, scoretanh=False, singlefc=True),
            "nocross_twofc": dict(gradkernels=True, maxngram=2, crossmatch=False, filters=48, scoretanh=False, singlefc=False),
            "uni_tanh": dict(gradkernels=False, maxngram=1, crossmatch=True, filters=128, scoretanh=True, singlefc=True),
        }.items():
            torch.manual_seed(500)
            rr = ref.ConvKNRM.ConvKNRM(cfg, provide={"extractor": ext})
            model = rr.build_model().eval()
            with torch.no_grad():
                # untrained features are O(100) per kernel x 9 views: scale the first combine layer so that tanh variants
                # do not saturate and the default score is O(1..10)
                model.combine[0].weight.mul_(0.05)
                if variant == "default" and shape_name != "full":
                    for i, k in enumerate(model.kernels.kernels):
                        k.mu.add_(0.011 * (i - 4))
                        k.sigma.mul_(1.0 + 0.04 * i)
                pos, neg = rr.score(tb)
                assert torch.equal(rr.test(tb), pos)
                if variant == "default":
                    # the [B,99] tensor fed to `combine`: recompute it with a hook on the combine layer's input
                    grabbed = []
                    h = model.combine[0].register_forward_hook(lambda m, i, o: grabbed.append(i[0].detach().clone()))
                    rr.test(tb)
                    h.remove()
                    out["feats"] = grabbed[0].numpy()
            out[f"{variant}/pos"] = pos.numpy()
            out[f"{variant}/neg"] = neg.numpy()
            out.update({f"{variant}/{k}": v for k, v in _state_np(model, skip=("embeddings.weight",)).items()})
        np.savez_compressed(GOLDEN / f"convknrm_{shape_name}.npz", **out)
        print("convknrm", shape_name, out["default/pos"][:4])


BERT_CONFIGS = {
    # name: (BertConfig kwargs, N docs, P passages, L, qlen, weight seed, input seed)
    "tiny": (dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=1000,
                  max_position_embeddings=64), 6, 3, 48, 6, 0, 3),
    "mid": (dict(hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=1024, vocab_size=5000,
                 max_position_embeddings=512, initializer_range=0.08), 4, 2, 200, 16, 0, 5),
    "base": (dict(), 4, 1, 512, 32, 0, 3),  # BERT-base: BASELINE.json configs[3] (monoBERT: P=1)
    # round 2: enough BERT-base sequences to MEASURE the error distribution of the bf16x3 engine (max and 99th percentile), not only
    # pass / fail on four of them: 64 ragged monoBERT sequences, and a BERT-MaxP case with P=4 passages per document
    "base64": (dict(), 64, 1, 512, 32, 0, 11, ["max"]),
    "base_p4": (dict(), 12, 4, 320, 24, 0, 12, ["max", "avg"]),
}


def bert_weight_checksum(model) -> np.ndarray:
    tot = 0.0
    for v in model.state_dict().values():
        if v.dtype.is_floating_point:
            tot += float(v.double().abs().sum())
    return np.array([tot])


def make_bert():
    mod = refshim.load_bertmaxp()
    only = [a for a in sys.argv[2:] if a in BERT_CONFIGS] if len(sys.argv) > 2 and sys.argv[1] == "bert" else list(BERT_CONFIGS)
    for name, spec in BERT_CONFIGS.items():
        if name not in only:
            continue
        cfg, N, P, L, qlen, wseed, iseed = spec[:7]
        aggs = spec[7] if len(spec) > 7 else ["max", "first", "sum", "avg"]
        import transformers

        vocab = transformers.BertConfig(**cfg).vocab_size
        batch = synthetic.bert_batch(N, seqlen=L, qlen=qlen, vocab=vocab, seed=iseed, numpassages=P)
        tb = _t(batch)
        out = {k: v.astype(np.int32) for k, v in batch.items()}
        out.update(reference_commit=np.array(refshim.REFERENCE_COMMIT), weight_seed=np.array(wseed), input_seed=np.array(iseed),
                   shape=np.array([N, P, L, qlen]))
        ext = refshim.FakeExtractor(None, numpassages=P, maxseqlen=L)
        for agg in aggs:
            with refshim.patched_bert_from_pretrained(cfg, seed=wseed):
                rr = mod.PTBERTMaxP(dict(pretrained="bert-base-uncased", aggregation=agg, hidden_dropout_prob=0.1),
                                    provide={"extractor": ext})
                model = rr.build_model().eval()
            with torch.no_grad():
                out[f"{agg}/scores"] = rr.test(tb).numpy()
                if agg == "max":
                    flat = lambda t: t.reshape(N * P, L)
                    logits = model.bert(flat(tb["pos_bert_input"]), attention_mask=flat(tb["pos_mask"]),
                                        token_type_ids=flat(tb["pos_seg"]))[0]
                    out["logits"] = logits.numpy()
                    out["weight_checksum"] = bert_weight_checksum(model.bert)
                    out["config_json"] = np.array(model.bert.config.to_json_string())
        np.savez_compressed(GOLDEN / f"bert_{name}.npz", **out)
        print("bert", name, out["logits"][:3].tolist())


CEDR_CONFIGS = {
    # name: (BertConfig kwargs, N docs, P passages, L, maxqlen, weight seed, input seed, reranker config)
    "tiny": (dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=1000, max_position_embeddings=64),
             6, 3, 48, 6, 0, 13, dict(simmat_layers=[0, 1, 2], combine_hidden=16, cls="avg")),
    "mid": (dict(hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=1024, vocab_size=5000, max_position_embeddings=512,
                 initializer_range=0.08), 4, 2, 200, 16, 0, 15, dict(simmat_layers=[1, 3], combine_hidden=0, cls="max")),
    "base": (dict(), 2, 2, 512, 32, 0, 17, dict(simmat_layers=list(range(13)), combine_hidden=128, cls="avg")),  # CEDRKNRM.py defaults (combine_hidden 128 keeps the fixture small), BERT-base
}


def make_cedrknrm():
    """the reference CEDRKNRM wrapper driving a seeded random-init HF BertModel."""
    mod = refshim.load_cedrknrm()
    for name, (cfg, N, P, L, maxqlen, wseed, iseed, rcfg) in CEDR_CONFIGS.items():
        import transformers

        vocab = transformers.BertConfig(**cfg).vocab_size
        batch = synthetic.cedr_batch(N, P, L, maxqlen, vocab=vocab, seed=iseed)
        tb = _t(batch)
        out = {k: v.astype(np.int32) for k, v in batch.items()}
        out.update(reference_commit=np.array(refshim.REFERENCE_COMMIT), weight_seed=np.array(wseed), input_seed=np.array(iseed),
                   shape=np.array([N, P, L, maxqlen]))
        ext = refshim.FakeExtractor(None, numpassages=P, maxseqlen=L, maxqlen=maxqlen)
        variants = {"default": rcfg, "nocls": {**rcfg, "cls": None}}
        if name == "tiny":
            variants["clsonly"] = {**rcfg, "simmat_layers": [-1]}
        for variant, vcfg in variants.items():
            full = dict(pretrained="bert-base-uncased", mus=[-0.9, -0.7, -0.5, -0.3, -0.1, 0.1, 0.3, 0.5, 0.7, 0.9], sigma=0.1, gradkernels=True,
                        hidden_dropout_prob=0.1, **vcfg)
            with refshim.patched_bertmodel_from_pretrained(cfg, seed=wseed):
                torch.manual_seed(600)
                rr = mod.CEDRKNRM(full, provide={"extractor": ext})
                model = rr.build_model().eval()
            with torch.no_grad():
                # the knrm features are 0.01*log sums (|x| ~ 1) against 768 cls dims: give them weight in the parity check
                model.combine[0].weight[:, -model.kernels.count() * len([l for l in vcfg["simmat_layers"] if l >= 0]):].mul_(5.0) if -1 not in vcfg["simmat_layers"] else None
                scores = rr.test(tb)
                out[f"{variant}/scores"] = scores.numpy()
                if variant == "default":
                    grabbed = []
                    h = model.combine[0].register_forward_hook(lambda m, i, o: grabbed.append(i[0].detach().clone()))
                    rr.test(tb)
                    h.remove()
                    out["feats"] = grabbed[0].numpy()
                    out["weight_checksum"] = bert_weight_checksum(model.bert)
                    out["config_json"] = np.array(model.bert.config.to_json_string())
            out.update({f"{variant}/{k}": v for k, v in _state_np(model, skip=("bert.", "one", "zero")).items()})
            out[f"{variant}/config_json"] = np.array(__import__("json").dumps(vcfg))
        np.savez_compressed(GOLDEN / f"cedrknrm_{name}.npz", **out)
        print("cedrknrm", name, out["default/scores"][:3].tolist())


PARADE_CONFIGS = {
    # name: (BertConfig kwargs, N docs, P passages, L, maxqlen, weight seed, input seed)
    "tiny": (dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, vocab_size=1000, max_position_embeddings=64),
             6, 3, 48, 6, 0, 23),
    # head dim 64 (the tensor-core attention path, also for the P+1 = 5-vector aggregator sequences); small enough to commit
    "mid": (dict(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, vocab_size=5000, max_position_embeddings=512,
                 initializer_range=0.08), 5, 4, 200, 16, 0, 25),
}


def make_parade():
    """the reference PTParade wrapper driving a seeded random-init HF BertModel.  (No BERT-base fixture:
    the two aggregation BertLayers alone are 14 M parameters that cannot be re-derived from a seed independently of torch's
    module-construction order; the BERT-base passage encoder itself is covered by the bert_base / cedrknrm_base goldens.)"""
    mod = refshim.load_parade()
    for name, (cfg, N, P, L, maxqlen, wseed, iseed) in PARADE_CONFIGS.items():
        import transformers

        vocab = transformers.BertConfig(**cfg).vocab_size
        batch = synthetic.cedr_batch(N, P, L, maxqlen, vocab=vocab, seed=iseed)
        tb = _t(batch)
        out = {k: v.astype(np.int32) for k, v in batch.items()}
        out.update(reference_commit=np.array(refshim.REFERENCE_COMMIT), weight_seed=np.array(wseed), input_seed=np.array(iseed),
                   shape=np.array([N, P, L, maxqlen]))
        ext = refshim.FakeExtractor(None, numpassages=P, maxseqlen=L, maxqlen=maxqlen)
        with refshim.patched_bertmodel_from_pretrained(cfg, seed=wseed):
            rr = mod.PTParade(dict(pretrained="bert-base-uncased", aggregation="transformer"), provide={"extractor": ext})
            model = rr.build_model().eval()
        with torch.no_grad():
            model.linear.weight.mul_(4.0)
            out["scores"] = rr.test(tb).numpy()
            flat = lambda t: t.reshape(N * P, L)
            cls = model.bert(flat(tb["pos_bert_input"]), attention_mask=flat(tb["pos_mask"]), token_type_ids=flat(tb["pos_seg"]))[0][:, 0, :]
            out["cls"] = cls.numpy()
            out["aggregated"] = model.aggregation(cls).numpy()
        out["weight_checksum"] = bert_weight_checksum(model.bert)
        out["config_json"] = np.array(model.bert.config.to_json_string())
        out.update(_state_np(model, skip=("bert.",)))
        np.savez_compressed(GOLDEN / f"parade_{name}.npz", **out)
        print("parade", name, out["scores"][:3].tolist())


TRAIN = dict(batch=32, itersize=512, niters=2, lr=1e-3, seed=4)


def make_knrm_train():
    """BASELINE.json configs[4]: reference PytorchTrainer.single_train_iteration (trainer/pytorch.py:76-122)
    driving the reference KNRM with Adam + pair_hinge_loss for niters=2."""
    ref = refshim.load_rerankers()
    tr = refshim.load_trainer()
    out = {}
    # Two well-posed settings (DESIGN.md "Exact matches": with gradkernels=True AND identical tokens in query and doc the
    # reference's own d/dmu, d/dsigma of the sigma=0.001 kernel is fp32 rounding noise that Adam turns into +-lr steps):
    #   frozen   : zipf triples with exact matches, gradkernels=False (only `combine` trains)
    #   disjoint : triples without shared terms, gradkernels=True (all 24 scalars train)
    # and, round 2, the reference DEFAULT as it is, so that the distance to it is a measured number:
    #   zipfgrad : zipf triples with exact matches, gradkernels=True (kernels.10.{mu,sigma} random-walk on that noise in the reference)
    for shape_name, setting in [(s_, t_) for s_ in ["full", "small"] for t_ in ["frozen", "disjoint", "zipfgrad"]]:
        B, Q, D, V, E, tseed, _ = SHAPES[shape_name]
        table = synthetic.embedding_table(V, E, seed=tseed)
        n_triples = TRAIN["itersize"] * TRAIN["niters"]
        data = synthetic.train_triples(n_triples, Q, D, V, seed=TRAIN["seed"], disjoint=setting == "disjoint")
        ext = refshim.FakeExtractor(table, maxqlen=Q, maxdoclen=D)
        torch.manual_seed(100)
        rr = ref.KNRM.KNRM(dict(gradkernels=setting != "frozen", scoretanh=False, singlefc=True, finetune=False), provide={"extractor": ext})
        shape_name = f"{shape_name}_{setting}"
        model = rr.build_model()
        with torch.no_grad():
            # untrained KNRM features are O(100); scale the combine layer so the hinge is active but not saturated
            model.combine[0].weight.mul_(0.02)
        init_state = {f"{shape_name}/init/{k[6:]}": v for k, v in _state_np(model).items()}
        trainer = tr.PytorchTrainer.__new__(tr.PytorchTrainer)
        trainer.config = dict(batch=TRAIN["batch"], evalbatch=0, niters=TRAIN["niters"], itersize=TRAIN["itersize"], gradacc=1,
                              lr=TRAIN["lr"], softmaxloss=False, fastforward=False, validatefreq=1, multithread=False,
                              boardname="default", warmupiters=0, decay=0.0, decayiters=3, decaytype=None, amp=None, seed=0)
        trainer.build()
        trainer.device = torch.device("cpu")
        model.train()
        trainer.optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=TRAIN["lr"])
        trainer.amp_train_autocast = contextlib.nullcontext
        trainer.scaler = None
        trainer.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(trainer.optimizer, trainer.lr_multiplier)
        trainer.loss = ref.common.pair_hinge_loss

        def batches():
            for s in range(0, n_triples, TRAIN["batch"]):
                yield {k: torch.from_numpy(v[s:s + TRAIN["batch"]]) for k, v in data.items()}

        it = batches()
        losses = []
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for niter in range(TRAIN["niters"]):
                losses.append(float(trainer.single_train_iteration(rr, it, niter)))
        out.update(init_state)
        out[f"{shape_name}/losses"] = np.array(losses)
        out.update({f"{shape_name}/final/{k[6:]}": v for k, v in _state_np(model).items()})
        out[f"{shape_name}/data_checksum"] = np.array([int(data["query"].sum()), int(data["posdoc"].sum()), int(data["negdoc"].sum())])
        out[f"{shape_name}/table_checksum"] = table_checksum(table)
        print("train", shape_name, losses)
    out["train_config"] = np.array(repr(TRAIN))
    np.savez_compressed(GOLDEN / "knrm_train.npz", **out)


def make_losses():
    ref = refshim.load_rerankers()
    rng = np.random.default_rng(7)
    pos, neg = rng.standard_normal(33).astype(np.float32) * 2, rng.standard_normal(33).astype(np.float32) * 2
    tp, tn = torch.from_numpy(pos), torch.from_numpy(neg)
    np.savez_compressed(GOLDEN / "losses.npz", pos=pos, neg=neg,
                        hinge=ref.common.pair_hinge_loss([tp, tn]).numpy(), softmax=ref.common.pair_softmax_loss([tp, tn]).numpy())


ALL = {"knrm": make_knrm, "drmm": make_drmm, "pacrr": make_pacrr, "drmmtks": make_drmmtks, "convknrm": make_convknrm, "cedrknrm": make_cedrknrm, "parade": make_parade, "bert": make_bert, "train": make_knrm_train, "losses": make_losses}

if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name in ([a for a in sys.argv[1:] if a in ALL] or list(ALL)):  # extra arguments select sub-fixtures (python -m oracle.make_goldens bert base64)
        ALL[name]()

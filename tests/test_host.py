"""CPU: host-side logic -- the C-ABI library loads and exports every declared symbol, the drop-in modules keep the
reference's API surface / checkpoint format, sharding arithmetic, and the world_size-2 gloo gather."""
import os
import pickle
import re
import socket
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_state, load_golden

from capreolus_b200 import _lib, reranker as R, sharding, synthetic


class Extractor:
    def __init__(self, V=50, E=20, Q=4, D=12, seed=0):
        self.embeddings = synthetic.embedding_table(V, E, seed=seed)
        self.config = {"maxqlen": Q, "maxdoclen": D}


def _declared_symbols():
    """(product, debug-only) symbol sets of include/capr_b200.h: debug-only = declared inside `#ifdef CAPR_DEBUG_BUILD`."""
    header = (ROOT / "include" / "capr_b200.h").read_text()
    product, debug, in_dbg = set(), set(), False
    for line in header.splitlines():
        if line.startswith("#ifdef CAPR_DEBUG_BUILD"):
            in_dbg = True
        elif in_dbg and line.startswith("#endif"):
            in_dbg = False
        for name in re.findall(r"\b(capr_[a-z0-9_]+)\s*\(", line):
            (debug if in_dbg else product).add(name)
    return product, debug


def test_library_exports_every_symbol_the_header_declares():
    declared, debug_only = _declared_symbols()
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert debug_only == set(_lib.DEBUG_SIGNATURES), debug_only ^ set(_lib.DEBUG_SIGNATURES)
    lib = _lib.lib()  # raises if the .so is missing or lacks a symbol
    for name in declared:
        assert hasattr(lib, name)
    for name in debug_only:  # micro-benchmarks / test hooks / profiling switches stay out of the product library
        assert not hasattr(lib, name), f"{name} must not be exported by the product library"
    dbg = _lib.dbg_lib()
    for name in declared | debug_only:
        assert hasattr(dbg, name)
    assert lib.capr_abi_version() == 1
    assert lib.capr_table_pitch(300) == 304 and lib.capr_table_pitch(16) == 16 and lib.capr_table_pitch(1) == 16


def test_abi_argument_validation_without_a_gpu():
    lib = _lib.lib()
    # argument checks run before any CUDA call, so they are testable on a CPU-only box
    rc = lib.capr_table_prepare(None, 10, 8, None, 16, None)
    assert rc == _lib.BAD_POINTER and b"null" in lib.capr_last_error()
    rc = lib.capr_table_prepare(None, 0, 8, None, 16, None)
    assert rc == _lib.BAD_SHAPE
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = lib.capr_knrm_forward(8, 8, 4, 40, 16, 16, 10, 16, 8, 8, 11, 8, 8, 0, None, None, 0, 8, None, None, None)
    assert rc == _lib.UNSUPPORTED and b"maxqlen" in lib.capr_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NativeLibraryMissing, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected_not_silently_scored():
    rr = R.KNRM(provide={"extractor": Extractor()})
    rr.build_model().eval()
    batch = {"query": torch.ones(2, 4, dtype=torch.long), "posdoc": torch.ones(2, 12, dtype=torch.long), "query_idf": torch.zeros(2, 4)}
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        rr.test(batch)


@pytest.mark.parametrize("cls,golden,variants", [
    ("KNRM", "knrm_small", ["default", "twofc", "tanh"]),
    ("DRMM", "drmm_small", ["default", "nh", "ch_tv"]),
    ("PACRR", "pacrr_small", ["default", "noidf_tanh", "wide"]),
    ("DRMMTKS", "drmmtks_small", ["default", "k3", "k20"]),
    ("ConvKNRM", "convknrm_small", ["default", "nocross_twofc", "uni_tanh"]),
])
def test_state_dict_keys_and_shapes_match_the_reference(cls, golden, variants):
    cfgs = {
        "KNRM": {"default": {}, "twofc": {"singlefc": False}, "tanh": {"scoretanh": True}},
        "DRMM": {"default": {}, "nh": {"histType": "NH"}, "ch_tv": {"nbins": 11, "nodes": 7, "histType": "CH", "gateType": "TV"}},
        "PACRR": {"default": {}, "noidf_tanh": {"idf": False, "nonlinearity": "tanh"},
                  "wide": {"mingram": 2, "nfilters": 16, "kmax": 3, "combine": 24, "nonlinearity": "none"}},
        "DRMMTKS": {"default": {}, "k3": {"topk": 3}, "k20": {"topk": 20, "freezeemb": False}},
        "ConvKNRM": {"default": {}, "nocross_twofc": {"maxngram": 2, "crossmatch": False, "filters": 48, "singlefc": False},
                     "uni_tanh": {"gradkernels": False, "maxngram": 1, "scoretanh": True}},
    }[cls]
    g = load_golden(golden)
    B, Q, D, V, E = (int(x) for x in g["shape"])
    for variant in variants:
        ref_state = golden_state(g, variant)
        model = getattr(R, cls)(cfgs[variant], provide={"extractor": Extractor(V, E, Q, D, seed=int(g["table_seed"]))}).build_model()
        ours = model.state_dict()
        assert set(ref_state) <= set(ours), set(ref_state) - set(ours)
        assert set(ours) - set(ref_state) <= {"embedding.weight", "simmat.embedding.weight", "embeddings.weight"}
        for k, v in ref_state.items():
            assert tuple(ours[k].shape) == tuple(v.shape), k
        model.load_state_dict(ref_state, strict=False)


def test_save_and_load_weights_use_the_reference_checkpoint_format(tmp_path):
    rr = R.KNRM(provide={"extractor": Extractor()})
    model = rr.build_model()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3)
    fn = tmp_path / "weights" / "dev.best"
    rr.save_weights(fn, opt)
    saved = pickle.load(open(fn, "rb"))
    assert not any("embedding.weight" in k or "_nosave_" in k for k in saved)  # reranker/__init__.py:34
    assert "kernels.kernels.10.sigma" in saved and "combine.0.weight" in saved
    assert Path(fn.as_posix() + ".optimizer").exists()
    with torch.no_grad():
        model.combine[0].weight.zero_()
    rr.load_weights(fn, opt)
    assert torch.equal(model.combine[0].weight, saved["combine.0.weight"])
    del saved["combine.0.bias"]
    pickle.dump(saved, open(fn, "wb"))
    with pytest.raises(RuntimeError, match="do not match"):
        rr.load_weights(fn, opt)


def test_config_handling_and_registry():
    rr = R.Reranker.create("DRMM", {"nbins": 11}, provide={"extractor": Extractor()})
    assert isinstance(rr, R.DRMM) and rr.config["nbins"] == 11 and rr.config["histType"] == "LCH"
    with pytest.raises(ValueError):
        R.KNRM({"nosuchoption": 1}, provide={"extractor": Extractor()})
    with pytest.raises(ValueError):
        R.Reranker.create("nosuchmodel")
    with pytest.raises(ValueError, match="gateType"):
        R.DRMM({"gateType": "XX"}, provide={"extractor": Extractor()}).build_model()
    m1 = rr.build_model()
    assert rr.build_model() is m1  # build_model is idempotent (DRMM.py:135-139)


def test_synthetic_inputs_are_deterministic_and_well_formed():
    a, b = synthetic.parity_batch(16, 8, 40, 500, seed=3), synthetic.parity_batch(16, 8, 40, 500, seed=3)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["query"].dtype == np.int64 and a["posdoc"].shape == (16, 40) and a["query_idf"].dtype == np.float32
    assert (a["query"] < 0).any() and (a["query"] == 0).any() and (a["query"][3] == 0).all()
    assert not (synthetic.parity_batch(16, 8, 40, 500, seed=3, oov=False)["query"] < 0).any()
    t = synthetic.embedding_table(100, 8, seed=1)
    assert (t[0] == 0).all() and t.dtype == np.float32
    bb = synthetic.bert_batch(3, seqlen=64, qlen=8, vocab=2000, seed=1, numpassages=2)
    ids, mask, seg = bb["pos_bert_input"], bb["pos_mask"], bb["pos_seg"]
    assert ids.shape == (3, 2, 64) and (ids[:, :, 0] == synthetic.CLS).all() and (ids[:, :, 9] == synthetic.SEP).all()
    assert (seg[:, :, :10] == 0).all() and (seg[:, :, 10:] == 1).all()  # segment ids stay 1 through the padding
    assert ((ids != 0) == (mask == 1)).all()


def test_shard_bounds_partition_exactly():
    for n in (0, 1, 7, 64, 100_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharding.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                cover.extend(range(lo, hi)) if n < 1000 else cover.append((lo, hi))
            if n < 1000:
                assert cover == list(range(n))
            else:
                assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


class _FakeReranker:
    """Deterministic per-pair scorer on CPU tensors (stands in for the CUDA path in the gloo test)."""

    def test(self, batch):
        return (batch["query"].float().sum(dim=1) * 0.5 + batch["posdoc"].float().sum(dim=1) * 0.25).view(-1)


def _gloo_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        batch = {"query": torch.from_numpy(rng.integers(0, 100, (n, 4))), "posdoc": torch.from_numpy(rng.integers(0, 100, (n, 9))),
                 "qid": [str(i) for i in range(n)]}
        got = sharding.ShardedScorer(_FakeReranker()).test(batch)
        want = _FakeReranker().test(batch)
        q.put((rank, bool(torch.equal(got, want)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 1])
def test_sharded_scoring_with_gloo_world_size_2(n):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(ok and shape == (n,) for _, ok, shape in results), results


# ---- packed id store + TREC run writer (host logic) -------------------------------------------
def test_packed_id_store_layout_and_validation():
    from capreolus_b200.predict import PackedIdStore

    st = PackedIdStore.from_lists({"d1": [5, 6, 7], "d2": [], "d3": [9]}, idf={"d1": [0.5, 1.0, 1.5], "d2": [], "d3": [2.0]})
    assert st.names == ["d1", "d2", "d3"] and st.index["d3"] == 2 and len(st) == 3
    assert st.flat.dtype == torch.int32 and st.flat.tolist() == [5, 6, 7, 9]
    assert st.offsets.dtype == torch.int64 and st.offsets.tolist() == [0, 3, 3, 4]
    assert st.idf.tolist() == [0.5, 1.0, 1.5, 2.0]
    with pytest.raises(ValueError, match="duplicate"):
        PackedIdStore(["a", "a"], [1, 2], [0, 1, 2])
    with pytest.raises(ValueError, match="offsets"):
        PackedIdStore(["a", "b"], [1, 2], [0, 1])
    with pytest.raises(ValueError, match="non-decreasing"):
        PackedIdStore(["a", "b"], [1, 2], [0, 3, 2])
    with pytest.raises(ValueError, match="idf"):
        PackedIdStore(["a"], [1, 2], [0, 2], idf=[1.0])


@pytest.mark.parametrize("n", [0, 1, 5, 777, 6216, 100_000])
@pytest.mark.parametrize("ramp", [True, False])
def test_pipelined_predictor_schedule_covers_every_item_once(n, ramp):
    """Chunk schedule of the host predict loop (`predict.PipelinedPredictor.schedule`): contiguous, complete, bounded by
    `chunk`; with `ramp` the first chunk is at most chunk/8 so the only non-overlapped H2D copy is short."""
    from capreolus_b200.predict import PipelinedPredictor

    p = PipelinedPredictor.__new__(PipelinedPredictor)  # no CUDA stream needed for the schedule
    p.chunk, p.ramp = 6216, ramp
    spans = p.schedule(n)
    assert [lo for lo, _ in spans] == [0] * bool(spans) + [hi for _, hi in spans[:-1]]
    assert (spans[-1][1] if spans else 0) == n
    assert all(0 < hi - lo <= p.chunk for lo, hi in spans)
    if ramp and n > p.chunk:
        assert spans[0][1] - spans[0][0] == p.chunk // 8


@pytest.mark.parametrize("k", [1, 3, 10])
def test_emulated_topk_insert_equals_torch_topk(k):
    """The branch-free sorted insert of drmmtks.cu (tests/emulate.py) keeps exactly torch.topk's multiset, duplicates included."""
    from emulate import emulated_topk_insert

    rng = np.random.default_rng(k)
    for trial in range(20):
        n = int(rng.integers(k, 80))
        v = rng.standard_normal(n).astype(np.float32)
        v[rng.integers(0, n, size=n // 3)] = 0.0  # zeros of padded columns compete (DRMMTKS over the zero-padded matrix)
        if trial % 4 == 0:
            v[rng.integers(0, n, size=3)] = 1.0  # several exact matches
        want = torch.topk(torch.from_numpy(v), k).values.numpy()
        assert np.array_equal(emulated_topk_insert(v, k), want)


def test_emulated_drmm_counts_equal_the_reference_histogram():
    """The one-pass binning of drmm.cu (tests/emulate.py) equals the reference's `sim < ub_i` counting + differencing
    (`DRMM.py:55-70`) on cosines that include bin edges, exact matches, pads and values >= 1."""
    from emulate import emulated_drmm_counts

    nbins = 29
    rng = np.random.default_rng(7)
    ub = torch.linspace(-1, 1, nbins + 1)[1:]
    for trial in range(10):
        D = 96
        sim = torch.from_numpy(rng.uniform(-1.0, 1.0, size=D).astype(np.float32))
        sim[:nbins] = ub[:nbins]                                   # exactly on every upper bound
        sim[nbins:2 * nbins - 1] = torch.nextafter(ub[:nbins - 1], torch.tensor(-2.0))  # one ulp below the bounds
        sim[60] = 0.9995                                           # inside the exact slot, below 1.0
        dids = rng.integers(1, 1000, size=D)
        dids[70:76] = 0                                            # pads: no bin
        sim[70:76] = 0.0
        want = torch.zeros(nbins + 1)
        real = torch.from_numpy(dids != 0)
        masked = torch.where(real, sim, torch.full_like(sim, 1e7))  # DRMM.py:59
        cum = torch.stack([(masked < ub[i]).sum() for i in range(nbins)]).float()
        want[0] = cum[0]
        want[1:nbins] = cum[1:] - cum[:-1]
        want[nbins] = ((masked > 0.999) & (masked < 1.001)).sum()
        got = emulated_drmm_counts(sim.numpy(), dids, qid=5, nbins=nbins)
        assert np.array_equal(got, want.numpy().astype(np.int64)), trial


def test_write_trec_run_matches_the_reference_format(tmp_path):
    """capreolus/searcher/__init__.py:48-58: qids in int order, docs by score descending (stable), 'qid Q0 docid rank score capreolus'."""
    from capreolus_b200.predict import write_trec_run

    preds = {"10": {"a": 0.5, "b": 2.0, "c": 0.5}, "9": {"x": -1.25}}
    fn = tmp_path / "run.txt"
    write_trec_run(preds, fn)
    assert fn.read_text().splitlines() == ["9 Q0 x 1 -1.25 capreolus", "10 Q0 b 1 2.0 capreolus", "10 Q0 a 2 0.5 capreolus", "10 Q0 c 3 0.5 capreolus"]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_default_sequences_per_encoder_call(monkeypatch):
    """2 x SM count (296 on a B200; the same value without a device) unless CAPR_BERT_SEQS_PER_CALL overrides it."""
    from capreolus_b200.reranker.ptBERTMaxP import default_seqs_per_call

    monkeypatch.delenv("CAPR_BERT_SEQS_PER_CALL", raising=False)
    n = default_seqs_per_call()
    assert n > 0 and n % 2 == 0
    import torch

    if not torch.cuda.is_available():
        assert n == 296
    monkeypatch.setenv("CAPR_BERT_SEQS_PER_CALL", "37")
    assert default_seqs_per_call() == 37

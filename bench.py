#!/usr/bin/env python
"""bench.py -- query-doc pairs scored per second (|q|=32, |d|=512), BASELINE.json's metric.

    python bench.py --gpus 1 --steps 10 --warmup 3                # this repo's CUDA path (default workload: configs[1], KNRM 100k pairs,
                                                                  #   + a `secondary` block: monoBERT, the tensor-pipe half of the metric)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # one rank per GPU, weak scaling + NCCL score gather
    python bench.py --impl reference --steps 3 --warmup 1         # the reference's CPU PyTorch path (oracle port) on host cores
    python bench.py --model bert | drmm | pacrr | drmmtks | convknrm   # the other configs of BASELINE.json (not the headline line)
    python bench.py --model bert --pairs 125000 --steps 1         # configs[3] at full size under torchrun with 8 ranks: 1000 q x 1000 docs
    python bench.py --mode train                                  # configs[4]: KNRM pairwise-hinge training, ms / iteration

A "step" is one pass of the hot path over one batch of synthetic (query, doc) pairs per GPU: ``reranker.test(batch)``
-> C ABI -> fused kernel(s) (+ one all-gather of the scores when N > 1).  Prints ONE JSON line (rank 0).
DESIGN.md "Measurement" states the byte / FLOP model behind ``roofline``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from capreolus_b200 import synthetic  # noqa: E402

Q, D, V, E = synthetic.MAXQLEN, synthetic.MAXDOCLEN, synthetic.VOCAB, synthetic.EMB_DIM
# ids (32+512)*8 B + gathered rows 544*300*4 B + one fp32 score
ALGO_BYTES_PER_PAIR = (Q + D) * 8 + (Q + D) * E * 4 + 4
BERT_L = 512
ENCODERS = ("bert", "cedrknrm", "parade")  # models whose hot path is the BERT encoder (tensor-pipe roofline)
PARADE_P, PARADE_L = 4, 163  # PARADE: the 512-token document as 4 passages of 128 tokens: [CLS] q(32) [SEP] passage(128) [SEP]
# per layer 2*12*768^2*512 (Linear layers) + 2*2*512^2*768 (attention) = 8.05 GFLOP; x12 layers = 96.6 GFLOP
BERT_FLOPS_PER_PAIR = 12 * (2 * 12 * 768 * 768 * 512 + 2 * 2 * 512 * 512 * 768)
MODELS = {"knrm": "KNRM", "drmm": "DRMM", "pacrr": "PACRR", "bert": "PTBERTMaxP", "drmmtks": "DRMMTKS", "convknrm": "ConvKNRM", "cedrknrm": "CEDRKNRM", "parade": "PTParade"}
DEFAULT_PAIRS = {"knrm": 100_000, "drmm": 100_000, "pacrr": 100_000, "bert": 1184, "drmmtks": 100_000, "convknrm": 100_000, "cedrknrm": 512, "parade": 1024}
# pairs per H2D chunk of the end-to-end pipeline: multiples of the 148 SMs for the persistent one-CTA-per-SM kernels (no ragged last wave)
# (KNRM, same box, 10 steps: chunk 6 216 -> e2e 10.17 / 10.26 M pairs/s, 12 432 -> 10.48 / 10.53 M, 24 864 -> 10.60 M: fewer per-chunk launches and pipeline refills)
DEFAULT_CHUNK = {"knrm": 24_864, "drmm": 24_864, "pacrr": 12_432, "bert": 296, "drmmtks": 24_864, "convknrm": 12_500, "cedrknrm": 128, "parade": 256}
TOP_KERNEL = {"knrm": "knrm_tc_kernel", "drmm": "drmm_tc_kernel", "pacrr": "pacrr_tc_kernel", "bert": "gemm2_kernel<3> (+ attention_tc4_kernel)",
              "cedrknrm": "gemm2_kernel<3> (+ attention_tc4_kernel, cedr_pool_kernel)", "parade": "gemm2_kernel<3> (+ attention_tc4_kernel)", "drmmtks": "drmmtks_tc_kernel", "convknrm": "knrm_tc_kernel x 9 views (+ convknrm_reps_kernel)"}
ORACLE_FN = {"knrm": "knrm_forward", "drmm": "drmm_forward", "pacrr": "pacrr_forward", "drmmtks": "drmmtks_forward", "convknrm": "convknrm_forward"}
METRIC = f"query-doc pairs scored/sec (|q|={Q},|d|={D})"
# CPU arm: one "step" = CPU_STEP_CALLS forwards of CPU_BATCH pairs (KNRM family) -- the same sample in `cpu_baseline` and in `--impl reference`
CPU_BATCH, CPU_STEP_CALLS = 64, 8


class Extractor:
    def __init__(self, table=None, **config):
        self.embeddings = table
        self.config = config


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": float(d["hbm_gbs"]), "tensor": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Polls NVML for SM clocks and throttle reasons while the timed regions run."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.active, self.stop_flag, self.max_mhz = [], set(), False, False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            if self.active:
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    self.reasons |= {name for bit, name in self.REASONS.items() if mask & bit}
                except Exception:
                    pass
            time.sleep(0.02)

    def summary(self, reset=False):
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if reset:
            self.samples, self.reasons = [], set()
        return out


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def build_reranker(model_key, config=None):
    """Random-init reranker of BASELINE.json's architecture (no network for checkpoints): returns (reranker, torch module)."""
    from capreolus_b200 import reranker as R

    torch.manual_seed(0)
    if model_key == "parade":
        rr = R.PTParade(dict(pretrained={}), provide={"extractor": Extractor(numpassages=PARADE_P, maxseqlen=PARADE_L, maxqlen=Q)})
    elif model_key == "cedrknrm":
        rr = R.CEDRKNRM(dict(pretrained={}, simmat_layers="0..12,1", cls="avg"), provide={"extractor": Extractor(numpassages=1, maxseqlen=BERT_L, maxqlen=Q)})
    elif model_key == "bert":
        rr = R.PTBERTMaxP(dict(pretrained={}, aggregation="max", hidden_dropout_prob=0.1), provide={"extractor": Extractor(numpassages=1, maxseqlen=BERT_L)})
    else:
        rr = getattr(R, MODELS[model_key])(config or {}, provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), maxqlen=Q, maxdoclen=D)})
    return rr, rr.build_model().eval()


def host_batch(model_key, n, seed):
    if model_key == "parade":
        b = synthetic.bert_batch(n, seqlen=PARADE_L, qlen=Q, seed=seed, numpassages=PARADE_P, ragged=False)
    elif model_key in ENCODERS:
        b = synthetic.bert_batch(n, seqlen=BERT_L, qlen=Q, seed=seed, numpassages=1, ragged=False)
    else:
        b = synthetic.throughput_batch(n, Q, D, V, seed=seed)
        if os.environ.get("CAPR_BENCH_IDS") == "uniform":  # experiment: no hot rows (not the BASELINE workload)
            rng = np.random.default_rng(seed)
            b["query"] = rng.integers(1, V, size=b["query"].shape)
            b["posdoc"] = rng.integers(1, V, size=b["posdoc"].shape)
    return {k: torch.from_numpy(v) for k, v in b.items()}


def workload_text(model_key, n):
    """`config.workload` -- the same string in this arm and in `--impl reference` (the reference arm times a bounded sample of it)."""
    if model_key in ENCODERS:
        names = {"bert": "monoBERT", "cedrknrm": "CEDR-KNRM (13 similarity layers, cls=avg) on", "parade": f"PARADE (transformer aggregation, {PARADE_P} passages of L={PARADE_L}) on"}
        shape = (f"L={BERT_L} (|q|={Q}, doc truncated to {BERT_L - Q - 3})" if model_key != "parade"
                 else f"|q|={Q}, |d|=512 as {PARADE_P} passages of {PARADE_L - Q - 3} tokens")
        return (f"{names[model_key]} (BERT-base, random init) forward, {n} synthetic pairs per GPU per step, {shape}, bf16x3 parity mode; "
                f"BASELINE.json configs[3] (1000 q x 1000 docs = 125000 pairs per GPU on 8 GPUs){'' if n >= 125000 else ', bounded sample'}")
    return (f"{MODELS[model_key]} forward, {n} synthetic pairs per GPU per step (BASELINE.json configs[1]), |q|={Q} |d|={D} vocab={V} "
            f"emb={E}, zipf ids, random-init weights")


def config_block(model_key, n, world):
    """`config` of the JSON line -- identical in this arm and in `--impl reference` (which times a bounded sample of the same workload)."""
    if model_key in ENCODERS:
        l2 = "activations of one 296-sequence encoder call (~5 GB) exceed L2; weights (0.35 GB as bf16 hi/lo planes) stream from HBM/L2"
    else:
        l2 = "inputs larger than L2 (435 MB of ids per step); the 36 MB embedding table is L2-resident by design"
    return {"workload": workload_text(model_key, n), "pairs_per_gpu": n, "l2_policy": l2,
            "parallelism": f"pairs sharded over {world} GPU(s), one all-gather of scores per step" if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port / HF module): `cpu_baseline` of our line and the whole `--impl reference` arm
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(model_key, state):
    """One bounded step of the reference's CPU PyTorch path (oracle port with the reference's op sequence; for BERT the HF
    module the reference itself calls).  Returns (step, description): step() runs one STEP of the CPU sample and returns pairs scored."""
    from oracle import restated

    torch.set_num_threads(os.cpu_count() or 1)
    if model_key == "parade":
        b = host_batch("parade", 4, seed=3)

        def step():
            with torch.no_grad():
                restated.parade_forward(state, b["pos_bert_input"], b["pos_mask"], b["pos_seg"], 12)
            return 4

        return step, f"oracle/restated.parade_forward (reference op sequence: BERT-base over {PARADE_P} passages of L={PARADE_L} + 2 aggregation BertLayers), 4 documents per step"
    if model_key == "cedrknrm":
        b = host_batch("cedrknrm", 4, seed=3)

        def step():
            with torch.no_grad():
                restated.cedrknrm_forward(state, b["pos_bert_input"], b["pos_mask"], b["pos_seg"], 12, Q, list(range(13)), "avg", 1024)
            return 4

        return step, "oracle/restated.cedrknrm_forward (reference op sequence: BERT-base hidden states, 13 masked cosine matrices, kernel pooling), 4 pairs per step, L=512"
    if model_key == "bert":
        import transformers

        torch.manual_seed(0)
        hf = transformers.BertForSequenceClassification(transformers.BertConfig()).eval()
        b = host_batch("bert", 8, seed=3)
        flat = {k: v.reshape(8, BERT_L) for k, v in b.items()}

        def step():
            with torch.no_grad():
                logits = hf(flat["pos_bert_input"], attention_mask=flat["pos_mask"], token_type_ids=flat["pos_seg"])[0]
                restated.bert_maxp_aggregate(logits[:, 1].reshape(8, 1), b["pos_mask"], b["pos_seg"], "max")
            return 8

        return step, "HF BertForSequenceClassification(BertConfig()) random init (the module ptBERTMaxP.py:82 calls), 8 sequences of L=512 per step, eval/no_grad"
    fn = getattr(restated, ORACLE_FN[model_key])
    table = torch.from_numpy(synthetic.embedding_table(V, E, seed=0))
    t = host_batch(model_key, CPU_BATCH * CPU_STEP_CALLS, seed=2)

    def step():
        with torch.no_grad():
            for i in range(CPU_STEP_CALLS):
                sl = slice(i * CPU_BATCH, (i + 1) * CPU_BATCH)
                fn(state, table, t["posdoc"][sl], t["query"][sl], t["query_idf"][sl])
        return CPU_BATCH * CPU_STEP_CALLS

    return step, (f"oracle/restated.{fn.__name__} (reference op sequence), {CPU_STEP_CALLS} batches of {CPU_BATCH} = {CPU_BATCH * CPU_STEP_CALLS} pairs per step, "
                  f"|q|={Q} |d|={D} V={V} E={E}")


#: glibc allocator settings of the CPU arm.  The reference's forward allocates its [B,11,32,512] temporaries (2 x 46 MB per batch of
#: 64) afresh on every call; with glibc's defaults each of them is mmap()ed, page-faulted in and unmapped again, which costs the CPU
#: path 3.5x (measured on the B200 host: 3 167 -> 11 216 pairs/s).  Keeping freed memory in the heap removes that -- the CPU arm
#: gets the faster setting so that the baseline is the reference's arithmetic, not its page faults.
MALLOC_ENV = {"MALLOC_MMAP_THRESHOLD_": "4294967296", "MALLOC_TRIM_THRESHOLD_": "8589934592", "MALLOC_TOP_PAD_": "1073741824"}


def cpu_baseline(model_key, pairs):
    """`cpu_baseline` of our line = the reference arm itself (`bench.py --impl reference`, same code path, allocator settings and
    procedure) run as a subprocess on the host cores: the two numbers of one record come from the same measurement."""
    import subprocess

    encoder = model_key in ENCODERS
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--model", model_key, "--pairs", str(pairs),
           "--steps", "2" if encoder else "16", "--warmup", "1" if encoder else "3"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    if out.returncode != 0 or not lines:
        return {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port", "sample": f"reference arm failed: {out.stderr[-300:]}"}
    ref = json.loads(lines[-1])
    cb = dict(ref["cpu_baseline"])
    cb["sample"] = ref["sample"] + f"; {ref['steps']} timed steps after {ref['warmup']} warm-up steps; torch {torch.__version__} CPU fp32"
    return cb


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if any(os.environ.get(k) != v for k, v in MALLOC_ENV.items()) and not os.environ.get("CAPR_BENCH_NO_REEXEC"):
        os.environ.update(MALLOC_ENV)  # the allocator reads these at start-up: re-execute this very command with them set
        os.execv(sys.executable, [sys.executable] + sys.argv)
    rr, model = build_reranker(args.model) if args.model != "bert" else (None, None)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()} if model is not None else None
    step, what = cpu_reference_step(args.model, state)
    warmup = max(args.warmup, 1 if args.model in ENCODERS else 3)
    for _ in range(warmup):
        step()
    t0, pairs, per_step = time.perf_counter(), 0, []
    for _ in range(args.steps):
        s0 = time.perf_counter()
        pairs += step()
        per_step.append(time.perf_counter() - s0)
    total = time.perf_counter() - t0
    value = pairs / total
    per = pairs // args.steps
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.model, args.pairs, max(1, args.gpus)),
        "sample": f"each step is a bounded sample of that workload: {per} pairs, host CPU ({what})",
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port", "sample": what,
                         "best_step_pairs_per_s": per / min(per_step), "median_step_pairs_per_s": per / float(np.median(per_step)),
                         "allocator": "glibc malloc with " + " ".join(f"{k}={v}" for k, v in MALLOC_ENV.items()) + " (keeps the reference's 46 MB temporaries in the heap instead of mmap/page-fault/munmap per call; 3.5x faster than the defaults on this host)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        self.rank, self.world, self.local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)
        self.sampler = None

    def barrier(self):
        import torch.distributed as dist

        if self.world > 1 and dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize(self.dev)


def time_kernel_only(rr, gpu, steps, warmup):
    """ms per `rr.test(gpu)` (CUDA events on the launching stream), no collective: used for the before / after NCCL-init A/B."""
    with torch.no_grad():
        for _ in range(warmup):
            rr.test(gpu)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record()
            rr.test(gpu)
            b.record()
        torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def l2_gather_probe(dev):
    """Second roofline of the KNRM-family kernels: the L2 -> SM gather rate of a pure-gather micro-kernel with the product
    kernel's access pattern (zipf rows of the bf16 hi/lo table, 16-byte cp.async into a shared-memory ring, nothing else).  Lives in
    the debug library (csrc/bench/gather_bench.cu); returns None when that library or symbol is absent."""
    try:
        from capreolus_b200 import _lib

        dbg = _lib.dbg_lib()
        if not hasattr(dbg, "capr_debug_gather_bench"):
            return None
        rng = np.random.default_rng(7)
        pitch = dbg.capr_table_pitch_bf16(E)
        hi = torch.randn((V, pitch), device=dev).to(torch.bfloat16)
        lo = torch.randn((V, pitch), device=dev).to(torch.bfloat16)
        n_rows = 148 * 128 * 256
        out = {}
        for pattern in ("zipf", "uniform"):
            ids = synthetic.zipf_ids(rng, (n_rows,), V) if pattern == "zipf" else rng.integers(1, V, size=n_rows)
            rows = torch.from_numpy(ids.astype(np.int32)).to(dev)
            best = None
            for stages in (6, 10):
                ms = []
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(dbg.capr_debug_gather_bench(hi.data_ptr(), lo.data_ptr(), V, pitch, rows.data_ptr(), n_rows, stages,
                                                           torch.cuda.current_stream(dev).cuda_stream), dbg)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    ms.append(e0.elapsed_time(e1))
                gbs = n_rows * pitch * 2 * 2 / (min(ms[1:]) * 1e-3) / 1e9
                if best is None or gbs > best["gbs"]:
                    best = {"gbs": gbs, "stages_in_flight": stages, "kb_in_flight_per_sm": stages * 16}
            out[pattern] = best
        return {**out["zipf"], "gbs_uniform_rows": out["uniform"]["gbs"]}
    except Exception as exc:  # the probe is an extra; the bench line does not depend on it
        return {"error": str(exc)[:200]}


def measure(model_key, ctx, n, chunk, steps, warmup, packed=True, pre_nccl=None, rr_model=None, warmup_pairs=0, skip_e2e=False):
    """Device-resident + end-to-end timing of one model at this rank's share of the work.  Returns the rank-0 dict (None elsewhere)."""
    import torch.distributed as dist

    from capreolus_b200.predict import PinnedBatch, PipelinedPredictor
    from capreolus_b200.sharding import gather_scores

    rank, world, dev, sampler = ctx.rank, ctx.world, ctx.dev, ctx.sampler
    rr, model = rr_model if rr_model is not None else build_reranker(model_key)
    model.to(dev)
    host = host_batch(model_key, n, seed=2 + rank)
    pinned = PinnedBatch(host)  # ids narrowed to int16 / int32 where they fit (widened on the device by capr_widen_ids)
    gpu = {k: v.to(dev) for k, v in host.items()}
    n_total = n * world
    with torch.no_grad():
        if 0 < warmup_pairs < n:  # full-size single-step runs (configs[3]: 125 000 BERT pairs per GPU): warm the kernels up on a slice
            small = {k: v[:warmup_pairs] for k, v in gpu.items()}
            for _ in range(warmup):
                rr.test(small)
        else:
            for _ in range(warmup):
                s = rr.test(gpu)
                scores = gather_scores(s, n_total) if world > 1 else s
        # ---- device-resident timing: K steps, CUDA events on the launching stream, max over ranks --------------
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        sampler.active = True
        ev0.record()
        for i in range(steps):
            k_ev[i][0].record()
            s = rr.test(gpu)
            k_ev[i][1].record()
            scores = gather_scores(s, n_total) if world > 1 else s
        ev1.record()
        ctx.barrier()
        sampler.active = False
        clocks = sampler.summary(reset=True)
        elapsed_ms = ev0.elapsed_time(ev1)
        kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
        # ---- end to end: pinned host ids -> H2D -> score -> D2H of the scores, through the public predict API ------
        mine = scores[rank * n:(rank + 1) * n] if world > 1 else scores
        check = not os.environ.get("CAPR_BENCH_NOCHECK")  # profiling runs against the debug library with CAPR_*_DEBUG switches produce invalid scores
        e2e_ms, e2e_clocks = None, None
        if not skip_e2e:
            pred = PipelinedPredictor(rr, dev, chunk=chunk, ramp=model_key not in ENCODERS)  # encoder models: H2D is negligible, keep full chunks
            for _ in range(2):
                pred.predict(pinned)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.barrier()
            sampler.active = True
            e0.record()
            for _ in range(steps):
                out = pred.predict(pinned)
                if world > 1:
                    gather_scores(out.to(dev, non_blocking=True), n_total)
            e1.record()
            ctx.barrier()
            sampler.active = False
            e2e_clocks = sampler.summary(reset=True)
            e2e_ms = e0.elapsed_time(e1)
            assert not check or torch.equal(out.to(dev), mine), "pipelined predict != direct test"
        # ---- end to end from a packed, device-resident id store: only (query, doc) indices cross PCIe ----
        packed_ms = None
        if packed and model_key not in ENCODERS:
            from capreolus_b200.predict import PackedIdStore, PairAssembler, RunPredictor

            t = host
            names = list(range(n))
            qs = PackedIdStore(names, t["query"].reshape(-1).numpy(), np.arange(n + 1, dtype=np.int64) * Q, idf=t["query_idf"].reshape(-1).numpy())
            ds = PackedIdStore(names, t["posdoc"].reshape(-1).numpy(), np.arange(n + 1, dtype=np.int64) * D)
            rp = RunPredictor(PairAssembler(qs, ds, Q, D, dev), chunk=int(os.environ.get("CAPR_BENCH_PACKED_CHUNK", chunk if chunk >= 24_864 else 4 * chunk)))  # no bulk H2D to hide: fewer, larger chunks keep the host launch loop off the critical path
            idx = torch.arange(n, dtype=torch.int32).pin_memory()
            host_scores = torch.empty(n, dtype=torch.float32).pin_memory()
            for _ in range(2):
                host_scores.copy_(rp.score_indices(rr, idx, idx), non_blocking=True)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.barrier()
            p0.record()
            for _ in range(steps):
                sc = rp.score_indices(rr, idx, idx)
                host_scores.copy_(sc, non_blocking=True)
                if world > 1:
                    gather_scores(sc, n_total)
            p1.record()
            ctx.barrier()
            packed_ms = p0.elapsed_time(p1)
            assert not check or torch.equal(host_scores.to(dev), mine), "packed-store predict != direct test"
    # ---- per-rank numbers: the line reports the MAX (the contract) and min / median / max + the full list, so that a slow rank shows
    mine_stats = [elapsed_ms, e2e_ms or 0.0, kernel_ms, packed_ms or 0.0, pre_nccl or 0.0, float(clocks["sm_mhz"] or 0.0)]
    per_rank = None
    if world > 1:
        t = torch.tensor(mine_stats, device=dev, dtype=torch.float64)
        allr = torch.empty((world, len(mine_stats)), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, t)
        allr = allr.cpu().numpy()
        elapsed_ms, e2e_max, kernel_ms, pm = (float(allr[:, i].max()) for i in range(4))
        e2e_ms = e2e_max if e2e_ms is not None else None
        packed_ms = pm if packed_ms is not None else None
        spread = lambda col: {"min": float(allr[:, col].min()), "median": float(np.median(allr[:, col])), "max": float(allr[:, col].max()),
                              "per_rank": [round(float(x), 4) for x in allr[:, col]]}
        per_rank = {"kernel_ms": spread(2), "step_ms_total": spread(0), "e2e_ms_total": spread(1), "sm_mhz_median": spread(5),
                    "kernel_ms_before_nccl_init": spread(4) if pre_nccl else None,
                    "note": "value / ms_per_step / e2e use the MAX over ranks (timing contract); kernel_ms_before_nccl_init is the same launch timed "
                            "in this process before dist.init_process_group (A/B for the N>=2 slowdown the round-1 driver run saw)"}
    if rank != 0:
        return None
    pk = peaks()
    if model_key in ENCODERS:
        flops_pair = BERT_FLOPS_PER_PAIR
        if model_key == "parade":  # P passages of PARADE_L tokens (the 2 aggregation layers over P+1 vectors are < 0.1 %)
            flops_pair = PARADE_P * 12 * (2 * 12 * 768 * 768 * PARADE_L + 2 * 2 * PARADE_L * PARADE_L * 768)
        achieved = flops_pair * n / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "traffic": None,
                "peak_source": pk["src"] + " bf16_tflops_sustained (kernels timed inside a long step)", "kernel": TOP_KERNEL[model_key],
                "algorithmic_flops_per_pair": flops_pair, "issued_tensor_flops_per_pair": 3 * flops_pair, "issued_frac": 3 * achieved / pk["tensor"],
                "note": "achieved counts ALGORITHMIC flops over the whole forward; the bf16x3 parity mode issues 3 tensor-core products per "
                        "algorithmic flop (Linear layers and attention), so frac is capped at 0.33; issued_frac = 3 x frac", "forward_ms": kernel_ms, "pairs_per_forward": n}
        from capreolus_b200.reranker.ptBERTMaxP import default_seqs_per_call

        spc = default_seqs_per_call()  # sequences per encoder call
        chunks = (n + spc - 1) // spc
        launches = steps * ((2 + 12 * 7 + 1) * chunks if model_key == "bert" else (1 + 12 * 7 + 4) * chunks if model_key == "cedrknrm"
                            else (1 + 12 * 7 + 3 + 2 * 7) * ((n * PARADE_P + spc - 1) // spc))
    else:
        achieved = ALGO_BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
        traffic, tsrc = None, None
        tf = ROOT / "profiles" / f"{model_key}_dram_traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_pair", 0) * n or None
            tsrc = f"profiles/{tf.name}: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, scaled per pair (from profile, not measured in this run)"
        roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": pk["src"] + " hbm_gbs", "kernel": TOP_KERNEL[model_key], "kernel_ms_per_launch": kernel_ms,
                "algorithmic_bytes_per_pair": ALGO_BYTES_PER_PAIR, "pairs_per_launch": n,
                "note": "achieved = logical gather bytes (SURVEY.md 8d: ids + 544 gathered fp32 rows + score) / kernel time; the 36 MB table is "
                        "L2-resident, so the gather is L2->SM traffic and frac can exceed 1; `traffic` is the DRAM bytes ncu measured per launch"}
        if model_key == "pacrr":
            # the byte line is not what bounds PACRR: its 1x1 / 2x2 / 3x3 convolutions with 32 filters are 32 x (1 + 4 + 9) multiply-adds per cell
            # of the 32 x 512 cosine tile = 14.7 MFLOP per pair on the fp32 pipe (FFMA2), against 148 SMs x 128 FMA/clk at the sampled clock
            conv_flops = 2.0 * 32 * (1 + 4 + 9) * Q * D
            peak_fp32 = 148 * 128 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12
            got = conv_flops * n / (kernel_ms * 1e-3) / 1e12
            roof["fp32_pipe"] = {"achieved": got, "peak": peak_fp32, "unit": "TFLOP/s", "frac": got / peak_fp32, "conv_flops_per_pair": conv_flops,
                                 "note": "PACRR is bound by its fp32 convolutions (DESIGN.md section 3), not by bytes: this is the binding roofline; peak = 148 SMs x "
                                         "128 FMA/clk x 2 at the SM clock sampled in the timed region"}
        launches = steps * (1 if model_key != "convknrm" else 13 * ((n + 4095) // 4096))  # ConvKNRM: zero row, 2 rep kernels, 9 views, combine per chunk
    line = {
        "metric": METRIC,
        "value": n_total * steps / (elapsed_ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if model_key not in ENCODERS else "bf16x3 (fp32 accumulate)", "data": "synthetic",
        "config": config_block(model_key, n, world),
        "roofline": roof,
        "e2e": {"value": (n_total * steps / (e2e_ms * 1e-3)) if e2e_ms else None, "unit": "pairs/s", "h2d_bytes_per_step": pinned.bytes_per_item * n,
                "d2h_bytes_per_step": 4 * n, "api": "capreolus_b200.predict.PipelinedPredictor(reranker).predict(PinnedBatch(host batch))",
                "host_id_dtypes": {k: str(v.dtype).replace("torch.", "") for k, v in pinned.tensors.items()},
                "note": "token ids travel in the narrowest integer type that holds them (int16 for a 30k vocabulary; the reference ships int64) "
                        "and are widened on the device by capr_widen_ids", "clocks": e2e_clocks},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if per_rank is not None:
        line["ranks"] = per_rank
    if packed_ms is not None:
        line["e2e_packed"] = {"value": n_total * steps / (packed_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * n,
                              "api": "capreolus_b200.predict.RunPredictor.score_indices: pinned host (query, doc) int32 indices -> H2D -> capr_assemble_pairs "
                                     "from the device-resident packed id store -> score -> D2H (extra to `e2e`)"}
    return line


def run_train(args):
    """configs[4]: KNRM pairwise-hinge training, synthetic triples, niters=2 x (itersize 512 / batch 32 = 16 batches), 1 GPU --
    `PairwiseTrainer.single_train_iteration` (capreolus/trainer/pytorch.py:76-122 + Adam :205) on the CUDA path, and the same loop
    through the oracle port (reference forward ops + torch autograd + Adam) on the host cores."""
    from capreolus_b200.trainer import PairwiseTrainer
    from oracle import restated

    assert torch.cuda.is_available(), "bench.py --mode train needs a GPU"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    batch, itersize, niters = 32, 512, 2
    n_triples = itersize * niters
    data = synthetic.train_triples(n_triples, Q, D, V, seed=4)

    def batches():
        for s in range(0, n_triples, batch):
            yield {k: torch.from_numpy(v[s:s + batch]) for k, v in data.items()}

    def gpu_run():
        rr, model = build_reranker("knrm", {"gradkernels": True})
        with torch.no_grad():
            model.combine[0].weight.mul_(0.02)  # untrained KNRM features are O(100): keep the hinge active but not saturated (as oracle/make_goldens.py does)
        tr = PairwiseTrainer(batch=batch, itersize=itersize, lr=1e-3, device=dev)
        tr.prepare(rr)
        it, losses, times = batches(), [], []
        for i in range(niters):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            losses.append(float(tr.single_train_iteration(rr, it, i)))  # float() synchronises: the loss is read back like the reference logs it
            times.append(time.perf_counter() - t0)
        return losses, times

    gpu_run()  # warm-up (library load, table preparation, allocator)
    losses, times = gpu_run()

    def cpu_run():
        rr, model = build_reranker("knrm", {"gradkernels": True})
        with torch.no_grad():
            model.combine[0].weight.mul_(0.02)
        params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "embedding" not in k) for k, v in model.state_dict().items()}
        opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3)
        table = torch.from_numpy(synthetic.embedding_table(V, E, seed=0))
        torch.set_num_threads(os.cpu_count() or 1)
        it, losses, times = batches(), [], []
        for _ in range(niters):
            t0, acc = time.perf_counter(), []
            for _ in range(itersize // batch):
                b = next(it)
                pos = restated.knrm_forward(params, table, b["posdoc"], b["query"]).view(-1)
                neg = restated.knrm_forward(params, table, b["negdoc"], b["query"]).view(-1)
                loss = restated.pair_hinge_loss(pos, neg)
                loss.backward()
                opt.step()
                opt.zero_grad()
                acc.append(float(loss))
            losses.append(float(np.mean(acc)))
            times.append(time.perf_counter() - t0)
        return losses, times

    cpu_losses, cpu_times = cpu_run()
    ms = 1e3 * float(np.mean(times))
    line = {
        "metric": "KNRM pairwise-hinge training, triples/s (BASELINE.json configs[4]: batch 32, itersize 512, niters 2, Adam lr 1e-3)",
        "value": itersize / float(np.mean(times)), "unit": "triples/s", "n_gpus": 1, "steps": niters, "warmup": niters, "ms_per_step": ms,
        "ms_per_iteration": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"KNRM (gradkernels=True) training iteration = 16 batches x 32 triples (2 forwards + closed-form backward + Adam each), |q|={Q} |d|={D} V={V} E={E}, zipf triples",
                   "timing": "host wall clock around single_train_iteration incl. H2D of every batch and the loss read-back (the loop is host-driven, 16 optimizer steps)"},
        "losses": losses,
        "cpu_baseline": {"value": itersize / float(np.mean(cpu_times)), "unit": "triples/s", "ms_per_iteration": 1e3 * float(np.mean(cpu_times)), "cores": torch.get_num_threads(),
                         "kind": "port", "sample": "the same 2 iterations: oracle/restated.knrm_forward (reference op sequence) + torch autograd + Adam on the host cores",
                         "losses": cpu_losses},
        "losses_note": "a timing run, not a parity check: with gradkernels=True and zipf triples (shared terms) the gradients of the sigma=0.001 kernel are "
                       "fp32 rounding noise in the reference itself (DESIGN.md section 4), so the two arms' losses drift apart by a few percent after the first "
                       "Adam step; the parity tests (tests/test_gpu_train.py) compare every other parameter and the well-posed settings",
        "reference_note": "BASELINE.md §2 row 5: the reference PytorchTrainer took 2.2-4.2 s per iteration on the survey container's 8 vCPUs",
        "gpu_launches": niters * (itersize // batch) * 3,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="score", choices=["score", "train"])
    ap.add_argument("--model", default="knrm", choices=sorted(MODELS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: 100k for KNRM/DRMM/PACRR = configs[1], 1184 = 4 encoder calls of 296 sequences for BERT)")
    ap.add_argument("--chunk", type=int, default=0, help="pairs per H2D chunk of the end-to-end pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the monoBERT `secondary` block of the default (KNRM) line")
    ap.add_argument("--secondary-pairs", type=int, default=1184, help="monoBERT pairs per GPU per step in the `secondary` block")
    ap.add_argument("--warmup-pairs", type=int, default=0, help="run the warm-up steps on the first N pairs only (full-size single-step runs, e.g. configs[3])")
    ap.add_argument("--skip-e2e", action="store_true", help="device-resident timing only (profile runs of the full-size configs; the driver's default line keeps e2e)")
    args = ap.parse_args()
    args.pairs = args.pairs or DEFAULT_PAIRS[args.model]
    args.chunk = args.chunk or DEFAULT_CHUNK[args.model]
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    ctx = Ctx()
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(ctx.dev)
    ctx.sampler = ClockSampler(ctx.local)
    ctx.sampler.start()

    rr, model = build_reranker(args.model)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()} if args.model != "bert" else None
    pre_nccl = None
    if ctx.world > 1:
        # A/B for the N>=2 slowdown of round 1 (SCALE_r01: knrm_tc_kernel 9.77 ms at N=1, 10.8 ms at every N>=2): time the very same
        # launch in this process BEFORE the NCCL communicator exists, then again (inside `measure`) after.
        model.to(ctx.dev)
        probe_n = min(args.pairs, 29_600 if args.model not in ENCODERS else 128)
        probe = {k: v.to(ctx.dev) for k, v in host_batch(args.model, probe_n, seed=2 + ctx.rank).items()}
        pre_nccl = time_kernel_only(rr, probe, steps=5, warmup=3) * (args.pairs / probe_n)
        del probe
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)

    line = measure(args.model, ctx, args.pairs, args.chunk, args.steps, args.warmup, pre_nccl=pre_nccl, rr_model=(rr, model), warmup_pairs=args.warmup_pairs,
                   skip_e2e=args.skip_e2e, packed=not args.skip_e2e)
    if args.model == "knrm" and ctx.rank == 0 and ctx.world == 1 and not os.environ.get("CAPR_BENCH_NO_L2PROBE"):
        probe = l2_gather_probe(ctx.dev)
        if probe and "gbs" in probe:
            logical = line["roofline"]["achieved"]
            line["roofline"]["l2_gather"] = {
                "peak": probe["gbs"], "unit": "GB/s", "kb_in_flight_per_sm": probe["kb_in_flight_per_sm"],
                "how": "csrc/bench/gather_bench.cu (debug library): 4 producer warps per SM gather zipf rows of a bf16 hi/lo table with 16-byte cp.async "
                       "into a shared-memory ring, a consumer frees the stages; bytes moved / CUDA-event time, best of 3, measured in this run",
                "logical_frac": logical / probe["gbs"],
                "peak_uniform_rows": probe["gbs_uniform_rows"], "logical_frac_of_uniform": logical / probe["gbs_uniform_rows"],
                "note": "second roofline.  `peak` = a stripped producer (no MMA, no pooling) on zipf rows, the benchmark's id distribution: its .cg gathers "
                        "saturate on the L2 lines of the hottest rows (5-7 TB/s whatever the ring depth or the number of SMs past ~75).  `peak_uniform_rows` = "
                        "the same kernel on uniform rows: the SM-side limit of the cp.async path (~90 GB/s per SM).  The product kernel moves its logical "
                        "gather bytes at `logical_frac` of the first and `logical_frac_of_uniform` of the second; it is paced by its 96 KB of gathers in "
                        "flight per SM over the loaded L2 round trip (DESIGN.md section 3), not by either rate"}
        elif probe:
            line["roofline"]["l2_gather"] = probe

    # ---- secondary: monoBERT in the same process at the same N (the tensor-pipe half of BASELINE.json's metric) ----
    if args.model == "knrm" and not args.no_secondary:
        del rr, model
        torch.cuda.empty_cache()
        sec_steps = min(args.steps, 5)
        sec = measure("bert", ctx, args.secondary_pairs, DEFAULT_CHUNK["bert"], sec_steps, 3, packed=False)
        if ctx.rank == 0:
            line["secondary"] = {k: sec[k] for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype", "config", "roofline", "e2e", "gpu_launches", "clocks") if k in sec}
            line["secondary"]["model"] = "monoBERT"
            if "ranks" in sec:
                line["secondary"]["ranks"] = sec["ranks"]
            line["gpu_launches"] += sec["gpu_launches"]
    if ctx.rank == 0:
        if not args.no_cpu_baseline and ctx.world == 1:
            line["cpu_baseline"] = cpu_baseline(args.model, args.pairs)
        print(json.dumps(line), flush=True)
    ctx.sampler.stop_flag = True
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

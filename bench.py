#!/usr/bin/env python
"""bench.py -- query-doc pairs scored per second (|q|=32, |d|=512), BASELINE.json's metric.

    python bench.py --gpus 1 --steps 10 --warmup 3                # this repo's CUDA path (default workload: configs[1], KNRM 100k pairs)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # one rank per GPU, weak scaling + NCCL score gather
    python bench.py --impl reference --steps 3 --warmup 1         # the reference's CPU PyTorch path (oracle port) on host cores
    python bench.py --model bert | drmm | pacrr | drmmtks | convknrm   # the other configs of BASELINE.json (not the headline line)

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
DEFAULT_PAIRS = {"knrm": 100_000, "drmm": 100_000, "pacrr": 100_000, "bert": 1024, "drmmtks": 100_000, "convknrm": 100_000, "cedrknrm": 512, "parade": 1024}
# pairs per H2D chunk of the end-to-end pipeline: multiples of the 148 SMs for the persistent one-CTA-per-SM kernels (no ragged last wave)
DEFAULT_CHUNK = {"knrm": 6_216, "drmm": 12_432, "pacrr": 12_432, "bert": 256, "drmmtks": 12_432, "convknrm": 12_500, "cedrknrm": 128, "parade": 256}
TOP_KERNEL = {"knrm": "knrm_tc_kernel", "drmm": "drmm_tc_kernel", "pacrr": "pacrr_tc_kernel", "bert": "gemm2_kernel<3> (+ attention_tc2_kernel)",
              "cedrknrm": "gemm2_kernel<3> (+ attention_tc2_kernel, cedr_pool_kernel)", "parade": "gemm2_kernel<3> (+ attention_tc2_kernel)", "drmmtks": "drmmtks_tc_kernel", "convknrm": "knrm_tc_kernel x 9 views (+ convknrm_reps_kernel)"}
ORACLE_FN = {"knrm": "knrm_forward", "drmm": "drmm_forward", "pacrr": "pacrr_forward", "drmmtks": "drmmtks_forward", "convknrm": "convknrm_forward"}


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

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def build_reranker(model_key):
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
        rr = getattr(R, MODELS[model_key])({}, provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), maxqlen=Q, maxdoclen=D)})
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


def cpu_reference_step(model_key, state):
    """One bounded step of the reference's CPU PyTorch path (oracle port with the reference's op sequence; for BERT the HF
    module the reference itself calls).  Returns a closure f() -> pairs scored."""
    from oracle import restated

    torch.set_num_threads(os.cpu_count() or 1)
    if model_key == "parade":
        b = host_batch("parade", 4, seed=3)

        def step():
            with torch.no_grad():
                restated.parade_forward(state, b["pos_bert_input"], b["pos_mask"], b["pos_seg"], 12)
            return 4

        return step, f"oracle/restated.parade_forward (reference op sequence: BERT-base over {PARADE_P} passages of L={PARADE_L} + 2 aggregation BertLayers), B=4"
    if model_key == "cedrknrm":
        b = host_batch("cedrknrm", 4, seed=3)

        def step():
            with torch.no_grad():
                restated.cedrknrm_forward(state, b["pos_bert_input"], b["pos_mask"], b["pos_seg"], 12, Q, list(range(13)), "avg", 1024)
            return 4

        return step, "oracle/restated.cedrknrm_forward (reference op sequence: BERT-base hidden states, 13 masked cosine matrices, kernel pooling), B=4, L=512"
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

        return step, "HF BertForSequenceClassification(BertConfig()) random init, B=8, L=512, eval/no_grad (the module ptBERTMaxP.py:82 calls)"
    fn = getattr(restated, ORACLE_FN[model_key])
    table = torch.from_numpy(synthetic.embedding_table(V, E, seed=0))
    t = host_batch(model_key, 512, seed=2)
    pos = [0]

    def step():
        i = pos[0] % 8
        pos[0] += 1
        sl = slice(i * 64, (i + 1) * 64)
        with torch.no_grad():
            fn(state, table, t["posdoc"][sl], t["query"][sl], t["query_idf"][sl])
        return 64

    return step, f"oracle/restated.{fn.__name__} (reference op sequence), batches of 64, |q|={Q} |d|={D} V={V} E={E}"


def cpu_baseline(model_key, state, seconds=12.0, max_pairs=8192):
    step, what = cpu_reference_step(model_key, state)
    step()  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        done += step()
        el = time.perf_counter() - t0
        if el >= seconds or done >= max_pairs:
            break
    return {"value": done / el, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} pairs in {el:.1f} s: {what}; torch {torch.__version__} CPU fp32"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    rr, model = build_reranker(args.model) if args.model != "bert" else (None, None)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()} if model is not None else None
    step, what = cpu_reference_step(args.model, state)
    per_step_calls = 8 if args.model not in ENCODERS else (2 if args.model == "bert" else 4)  # 512 pairs (KNRM family) / 16 sequences (encoders) per step
    for _ in range(args.warmup):
        step()
    t0, pairs = time.perf_counter(), 0
    for _ in range(args.steps):
        for _ in range(per_step_calls):
            pairs += step()
    total = time.perf_counter() - t0
    value = pairs / total
    line = {
        "impl": "reference", "metric": f"query-doc pairs scored/sec (|q|={Q},|d|={D})", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODELS[args.model]} forward, {pairs // args.steps} synthetic pairs per step (a bounded sample of the "
                               f"{args.pairs}-pair workload), |q|={Q} |d|={D}, host CPU"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="knrm", choices=sorted(MODELS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: 100k for KNRM/DRMM/PACRR = configs[1], 1024 for BERT)")
    ap.add_argument("--chunk", type=int, default=0, help="pairs per H2D chunk of the end-to-end pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.pairs = args.pairs or DEFAULT_PAIRS[args.model]
    args.chunk = args.chunk or DEFAULT_CHUNK[args.model]
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    from capreolus_b200.predict import PinnedBatch, PipelinedPredictor
    from capreolus_b200.sharding import gather_scores

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    rr, model = build_reranker(args.model)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()} if args.model != "bert" else None
    model.to(dev)
    n = args.pairs
    pinned = PinnedBatch(host_batch(args.model, n, seed=2 + rank))
    gpu = {k: v.to(dev) for k, v in pinned.tensors.items()}
    n_total = n * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    with torch.no_grad():
        for _ in range(args.warmup):
            s = rr.test(gpu)
            scores = gather_scores(s, n_total) if world > 1 else s
        # ---- device-resident timing: K steps, CUDA events on the launching stream, max over ranks --------------
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.active = True
        ev0.record()
        for i in range(args.steps):
            k_ev[i][0].record()
            s = rr.test(gpu)
            k_ev[i][1].record()
            scores = gather_scores(s, n_total) if world > 1 else s
        ev1.record()
        barrier()
        sampler.active = False
        elapsed_ms = ev0.elapsed_time(ev1)
        kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
        # ---- end to end: pinned host ids -> H2D -> score -> D2H of the scores, through the public predict API ------
        pred = PipelinedPredictor(rr, dev, chunk=args.chunk, ramp=args.model not in ("bert", "cedrknrm", "parade"))  # encoder models: H2D is negligible, keep full chunks
        for _ in range(2):
            pred.predict(pinned)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.active = True
        e0.record()
        for _ in range(args.steps):
            out = pred.predict(pinned)
            if world > 1:
                gather_scores(out.to(dev, non_blocking=True), n_total)
        e1.record()
        barrier()
        sampler.active = False
        sampler.stop_flag = True
        e2e_ms = e0.elapsed_time(e1)
        mine = scores[rank * n:(rank + 1) * n] if world > 1 else scores
        check = not os.environ.get("CAPR_BENCH_NOCHECK")  # profiling runs with CAPR_*_DEBUG switches produce invalid scores
        assert not check or torch.equal(out.to(dev), mine), "pipelined predict != direct test"
        # ---- end to end from a packed, device-resident id store: only (query, doc) indices cross PCIe ----
        packed_ms = None
        if args.model not in ENCODERS:
            from capreolus_b200.predict import PackedIdStore, PairAssembler, RunPredictor

            t = pinned.tensors
            names = list(range(n))
            qs = PackedIdStore(names, t["query"].reshape(-1).numpy(), np.arange(n + 1, dtype=np.int64) * Q, idf=t["query_idf"].reshape(-1).numpy())
            ds = PackedIdStore(names, t["posdoc"].reshape(-1).numpy(), np.arange(n + 1, dtype=np.int64) * D)
            rp = RunPredictor(PairAssembler(qs, ds, Q, D, dev), chunk=int(os.environ.get("CAPR_BENCH_PACKED_CHUNK", 4 * args.chunk)))  # no bulk H2D to hide: fewer, larger chunks keep the host launch loop off the critical path
            idx = torch.arange(n, dtype=torch.int32).pin_memory()
            host_scores = torch.empty(n, dtype=torch.float32).pin_memory()
            for _ in range(2):
                host_scores.copy_(rp.score_indices(rr, idx, idx), non_blocking=True)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            p0.record()
            for _ in range(args.steps):
                sc = rp.score_indices(rr, idx, idx)
                host_scores.copy_(sc, non_blocking=True)
                if world > 1:
                    gather_scores(sc, n_total)
            p1.record()
            barrier()
            packed_ms = p0.elapsed_time(p1)
            assert not check or torch.equal(host_scores.to(dev), mine), "packed-store predict != direct test"
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, kernel_ms, packed_ms or 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, kernel_ms, pm = (float(x) for x in t)
        packed_ms = pm if packed_ms is not None else None

    if rank == 0:
        pk = peaks()
        if args.model in ENCODERS:
            flops_pair = BERT_FLOPS_PER_PAIR
            if args.model == "parade":  # P passages of PARADE_L tokens (the 2 aggregation layers over P+1 vectors are < 0.1 %)
                flops_pair = PARADE_P * 12 * (2 * 12 * 768 * 768 * PARADE_L + 2 * 2 * PARADE_L * PARADE_L * 768)
            achieved = flops_pair * n / (kernel_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "traffic": None,
                    "peak_source": pk["src"] + " bf16_tflops_sustained (kernels timed inside a long step)", "kernel": TOP_KERNEL[args.model],
                    "algorithmic_flops_per_pair": flops_pair, "issued_tensor_flops_per_pair": 3 * flops_pair,
                    "note": "achieved counts ALGORITHMIC flops over the whole forward; the bf16x3 parity mode issues 3 tensor-core products per "
                            "algorithmic flop (Linear layers and attention)", "forward_ms": kernel_ms, "pairs_per_forward": n}
            launches = args.steps * ((2 + 12 * 7) * ((n + 127) // 128) if args.model == "bert" else (1 + 12 * 7 + 4) * ((n + 127) // 128) if args.model == "cedrknrm"
                                    else (1 + 12 * 7 + 3 + 2 * 7) * ((n * PARADE_P + 127) // 128))
            names = {"bert": "monoBERT", "cedrknrm": "CEDR-KNRM (13 similarity layers, cls=avg) on", "parade": f"PARADE (transformer aggregation, {PARADE_P} passages of L={PARADE_L}) on"}
            shape = (f"L={BERT_L} (|q|={Q}, doc truncated to {BERT_L - Q - 3})" if args.model != "parade"
                     else f"|q|={Q}, |d|=512 as {PARADE_P} passages of {PARADE_L - Q - 3} tokens")
            workload = (f"{names[args.model]} (BERT-base, random init) forward, {n} synthetic pairs per GPU per step, {shape}, bf16x3 parity mode; "
                        f"bounded sample of BASELINE.json configs[3] (1000 q x 1000 docs)")
            l2 = "activations of one 128-sequence chunk (~2 GB) exceed L2; weights (0.35 GB as bf16 hi/lo planes) stream from HBM/L2"
        else:
            achieved = ALGO_BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
            traffic = None
            tf = ROOT / "profiles" / f"{args.model}_dram_traffic.json"
            if tf.exists():
                traffic = json.loads(tf.read_text()).get("dram_bytes_per_pair", 0) * n or None
            roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "traffic": traffic,
                    "peak_source": pk["src"] + " hbm_gbs", "kernel": TOP_KERNEL[args.model], "kernel_ms_per_launch": kernel_ms,
                    "algorithmic_bytes_per_pair": ALGO_BYTES_PER_PAIR, "pairs_per_launch": n,
                    "note": "achieved = logical gather bytes (SURVEY.md 8d: ids + 544 gathered fp32 rows + score) / kernel time; the 36 MB table is "
                            "L2-resident, so the gather is L2->SM traffic and frac can exceed 1; `traffic` is the DRAM bytes ncu measured per launch"}
            launches = args.steps * (1 if args.model != "convknrm" else 13 * ((n + 4095) // 4096))  # ConvKNRM: zero row, 2 rep kernels, 9 views, combine per chunk
            workload = (f"{MODELS[args.model]} forward, {n} synthetic pairs per GPU per step (BASELINE.json configs[1]), |q|={Q} |d|={D} vocab={V} "
                        f"emb={E}, zipf ids, random-init weights")
            l2 = "inputs larger than L2 (435 MB of ids per step); the 36 MB embedding table is L2-resident by design"
        line = {
            "metric": f"query-doc pairs scored/sec (|q|={Q},|d|={D})",
            "value": n_total * args.steps / (elapsed_ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.model not in ENCODERS else "bf16x3 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": workload, "pairs_per_gpu": n, "l2_policy": l2,
                       "parallelism": f"pairs sharded over {world} GPU(s), one all-gather of scores per step" if world > 1 else "single GPU"},
            "roofline": roof,
            "e2e": {"value": n_total * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": pinned.bytes_per_item * n,
                    "d2h_bytes_per_step": 4 * n, "api": "capreolus_b200.predict.PipelinedPredictor(reranker).predict(pinned host batch)"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
        }
        if packed_ms is not None:
            line["e2e_packed"] = {"value": n_total * args.steps / (packed_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 4 * n,
                                  "api": "capreolus_b200.predict.RunPredictor.score_indices: pinned host (query, doc) int32 indices -> H2D -> capr_assemble_pairs "
                                         "from the device-resident packed id store -> score -> D2H (extra to `e2e`, which ships padded int64 ids)"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.model, state)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

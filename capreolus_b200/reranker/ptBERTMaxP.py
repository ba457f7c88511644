"""monoBERT / BERT-MaxP behind the reference's module API (``capreolus/reranker/ptBERTMaxP.py``).

``PTBERTMaxP_Class`` keeps the HF ``BertForSequenceClassification`` as ``self.bert`` exactly like the reference (so
``state_dict`` keys, ``save_weights`` / ``load_weights`` and ``from_pretrained`` behave the same); in eval mode its
``forward`` does not run the torch module but the sm_100a encoder behind ``capr_bert_forward`` (tcgen05 GEMMs), built
lazily from the module's current parameters.  Passage aggregation (max / first / sum / avg) is the reference's
(``ptBERTMaxP.py:85-94``), including ``avg`` dividing by the passage-mask sum of the whole batch.
"""
from __future__ import annotations

import ctypes
import os

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption, Dependency
from capreolus_b200.reranker import Reranker

_LAYER_KEYS = [
    "attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight", "attention.self.key.bias",
    "attention.self.value.weight", "attention.self.value.bias", "attention.output.dense.weight", "attention.output.dense.bias",
    "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
    "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias",
]


def bert_weight_keys(n_layers: int, prefix: str = "bert.", classifier: bool = True) -> list[str]:
    """state_dict keys of a HF BertForSequenceClassification (``prefix='bert.'``) or BertModel (``prefix=''``,
    ``classifier=False``) in the order ``capr_bert_create`` expects them."""
    keys = [prefix + "embeddings.word_embeddings.weight", prefix + "embeddings.position_embeddings.weight",
            prefix + "embeddings.token_type_embeddings.weight", prefix + "embeddings.LayerNorm.weight", prefix + "embeddings.LayerNorm.bias"]
    for i in range(n_layers):
        keys += [f"{prefix}encoder.layer.{i}.{k}" for k in _LAYER_KEYS]
    keys += [prefix + "pooler.dense.weight", prefix + "pooler.dense.bias"]
    return keys + (["classifier.weight", "classifier.bias"] if classifier else [])


def default_seqs_per_call() -> int:
    """Sequences per encoder call = twice the SM count (296 on a B200).  The GEMMs run on clusters of two CTAs over 256 x 256 output tiles
    and the attention on one persistent CTA per SM: with a multiple of 148 sequences of 512 tokens every GEMM of BERT-base has a tile
    count divisible by the 74 clusters and the attention a whole number of items per CTA -- no ragged last wave (128-sequence calls
    lose 5.6 % of the N = 768 GEMMs: 768 tiles = 10.4 waves).  Measured on one box (monoBERT, bf16x3, under sw_power_cap): 128 per call
    3 697 pairs/s, 148: 3 735, 296: 3 807; the workspace is 17 MB per sequence (5 GB at 296).  ``CAPR_BERT_SEQS_PER_CALL`` overrides."""
    env = os.environ.get("CAPR_BERT_SEQS_PER_CALL")
    if env:
        return max(1, int(env))
    try:
        sms = _lib.lib().capr_device_sm_count()
    except Exception:  # no library / no device yet (models are also constructed on CPU boxes): the B200 value
        sms = 0
    return 2 * (sms if sms > 0 else 148)


class BertEngine:
    """Owns one ``capr_bert_t`` handle (a snapshot of the weights as bf16 planes + TMA descriptors) and its workspace."""

    PRECISION = {"bf16x3": _lib.BERT_BF16X3, "bf16": _lib.BERT_BF16}

    def __init__(self, hf_model, precision="bf16x3", max_seqs_per_call=None):
        cfg = hf_model.config
        # Electra's encoder (CEDR-KNRM's default, CEDRKNRM.py:20-27) is the BERT encoder under other parameter names as long as
        # there is no embedding projection (embedding_size == hidden_size: electra-base); it has no pooler
        electra = cfg.model_type == "electra"
        if cfg.model_type not in ("bert", "electra"):
            raise ValueError(f"capreolus_b200 encoder engine supports BERT / Electra encoders only, got {cfg.model_type!r}")
        if electra and getattr(cfg, "embedding_size", cfg.hidden_size) != cfg.hidden_size:
            raise ValueError("capreolus_b200 encoder engine: Electra with an embedding projection (embedding_size != hidden_size) is not implemented")
        if getattr(cfg, "hidden_act", "gelu") != "gelu":
            raise ValueError("capreolus_b200 ptBERTMaxP: only hidden_act='gelu' (erf) is implemented")
        if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise ValueError("capreolus_b200 ptBERTMaxP: only absolute position embeddings are implemented")
        state = hf_model.state_dict()
        headless = "classifier.weight" not in state  # a bare BertModel (CEDR-KNRM): no classification head
        keys = bert_weight_keys(cfg.num_hidden_layers, "" if headless else "bert.", classifier=not headless)
        if electra and not headless:
            raise ValueError("capreolus_b200 encoder engine: Electra classification heads are not implemented (hidden states only)")
        some = state[keys[0]]
        _lib.require_cuda(some)
        H = cfg.hidden_size
        if electra:  # ElectraModel has no pooler: zeros (a headless engine never evaluates it)
            state = {**state, "pooler.dense.weight": torch.zeros((H, H), device=some.device), "pooler.dense.bias": torch.zeros(H, device=some.device)}
        tensors = [state[k].detach().float().contiguous() for k in keys]
        _lib.require_cuda(*tensors)
        self.device = tensors[0].device
        if headless:  # capr_bert_create wants a classifier: hand it zeros (logits are never requested from such an engine)
            tensors += [torch.zeros((2, cfg.hidden_size), device=self.device), torch.zeros(2, device=self.device)]
        self.headless = headless
        self.hidden_size, self.n_layers = cfg.hidden_size, cfg.num_hidden_layers
        self.n_labels = tensors[-2].shape[0]
        self.cfg = _lib.BertConfigStruct(cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size,
                                         cfg.vocab_size, cfg.max_position_embeddings, cfg.type_vocab_size, self.n_labels,
                                         float(cfg.layer_norm_eps))
        self._create(tensors, precision)
        self.max_seqs_per_call = int(max_seqs_per_call or default_seqs_per_call())

    def _create(self, tensors, precision):
        ptrs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        handle = ctypes.c_void_p()
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(lib.capr_bert_create(ctypes.byref(self.cfg), ptrs, len(tensors), self.PRECISION[precision],
                                            _lib.current_stream(self.device), ctypes.byref(handle)))
            torch.cuda.current_stream(self.device).synchronize()  # the snapshot copies read `tensors`
        self.handle = handle
        self._workspace = None

    @classmethod
    def from_layers(cls, bert_layers, hf_config, precision="bf16x3"):
        """An engine made of stand-alone HF ``BertLayer`` modules (PARADE's ``transformer_layer_1/2``, ptparade.py:27-28): only its
        encoder layers are ever run (``capr_parade_head``); the embedding / pooler / classifier slots get zeros."""
        self = cls.__new__(cls)
        H = hf_config.hidden_size
        layer_tensors = []
        for layer in bert_layers:
            st = layer.state_dict()
            layer_tensors += [st[k].detach().float().contiguous() for k in _LAYER_KEYS]
        _lib.require_cuda(*layer_tensors)
        self.device = layer_tensors[0].device
        z = lambda *shape: torch.zeros(shape, device=self.device)
        tensors = [z(2, H), z(2, H), z(2, H), z(H), z(H)] + layer_tensors + [z(H, H), z(H), z(2, H), z(2)]
        self.headless, self.hidden_size, self.n_layers, self.n_labels = True, H, len(bert_layers), 2
        self.cfg = _lib.BertConfigStruct(H, len(bert_layers), hf_config.num_attention_heads, hf_config.intermediate_size, 2, 2, 2, 2,
                                         float(hf_config.layer_norm_eps))
        self._create(tensors, precision)
        self.max_seqs_per_call = 0
        return self

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().capr_bert_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _ws(self, n, L, device):
        need = _lib.lib().capr_bert_workspace_bytes(self.handle, n, L)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
        return self._workspace

    def hidden_states(self, ids, mask, seg, layers) -> torch.Tensor:
        """``[N,L]`` int64 x3 -> fp32 ``[len(layers), N*L, H]``: HF ``hidden_states[l]`` for l in ``layers`` (0 = embedding
        output).  One call: the caller chunks N (``capr_bert_forward_hidden``)."""
        _lib.require_cuda(ids, mask, seg)
        ids, mask, seg = (t.long().contiguous() for t in (ids, mask, seg))
        N, L = ids.shape
        layers = [int(l) for l in layers]
        out = torch.empty((len(layers), N * L, self.hidden_size), dtype=torch.float32, device=ids.device)
        with torch.cuda.device(ids.device):
            ws = self._ws(N, L, ids.device)
            arr = (ctypes.c_int * len(layers))(*layers)
            _lib.check(_lib.lib().capr_bert_forward_hidden(self.handle, ids.data_ptr(), mask.data_ptr(), seg.data_ptr(), N, L, arr, len(layers),
                                                           out.data_ptr(), None, ws.data_ptr(), ws.numel(), _lib.current_stream(ids.device)))
        return out

    def logits(self, ids, mask, seg) -> torch.Tensor:
        """``[N,L]`` int64 x3 -> classifier logits ``[N,n_labels]`` fp32."""
        _lib.require_cuda(ids, mask, seg)
        lib = _lib.lib()
        ids, mask, seg = (t.long().contiguous() for t in (ids, mask, seg))
        N, L = ids.shape
        out = torch.empty((N, self.n_labels), dtype=torch.float32, device=ids.device)
        step = max(1, self.max_seqs_per_call)
        with torch.cuda.device(ids.device):
            for lo in range(0, N, step):
                n = min(step, N - lo)
                need = lib.capr_bert_workspace_bytes(self.handle, n, L)
                if self._workspace is None or self._workspace.numel() < need or self._workspace.device != ids.device:
                    self._workspace = torch.empty(need, dtype=torch.uint8, device=ids.device)
                _lib.check(lib.capr_bert_forward(self.handle, ids[lo:lo + n].data_ptr(), mask[lo:lo + n].data_ptr(), seg[lo:lo + n].data_ptr(), n, L,
                                                 out[lo:lo + n].data_ptr(), self._workspace.data_ptr(), self._workspace.numel(),
                                                 _lib.current_stream(ids.device)))
        return out


class PTBERTMaxP_Class(nn.Module):
    """``PTBERTMaxP_Class`` (capreolus/reranker/ptBERTMaxP.py:29-96)."""

    def __init__(self, extractor, config, *args, **kwargs):
        super(PTBERTMaxP_Class, self).__init__(*args, **kwargs)
        import transformers

        self.extractor = extractor
        pretrained = config["pretrained"]
        if isinstance(pretrained, dict):
            # offline extension: a BertConfig dict -> random-init encoder (there is no network for checkpoints here)
            self.bert = transformers.BertForSequenceClassification(
                transformers.BertConfig(**{**pretrained, "hidden_dropout_prob": config["hidden_dropout_prob"]}))
        elif "electra" in pretrained or "roberta" in pretrained:
            raise ValueError(f"capreolus_b200 ptBERTMaxP: {pretrained!r} is not a BERT encoder (Electra / RoBERTa variants are out of scope)")
        elif pretrained == "bert-base-msmarco":
            self.bert = transformers.AutoModelForSequenceClassification.from_pretrained("Capreolus/bert-base-msmarco")
        else:
            self.bert = transformers.AutoModelForSequenceClassification.from_pretrained(
                pretrained, hidden_dropout_prob=config["hidden_dropout_prob"])
        self.config = config
        self.precision = config.get("precision", "bf16x3") if hasattr(config, "get") else "bf16x3"
        self._engine, self._engine_key = None, None

    def engine(self) -> BertEngine:
        params = list(self.bert.parameters())
        key = (params[0].device, tuple(p._version for p in params), params[0].data_ptr())
        if self._engine is None or key != self._engine_key:
            self._engine = BertEngine(self.bert, self.precision)
            self._engine_key = key
        return self._engine

    def forward(self, doc_input, doc_mask, doc_seg):
        if self.training:
            raise NotImplementedError("capreolus_b200 ptBERTMaxP: only inference (model.eval()) is implemented; BERT training is out of scope")
        return self.predict_step(doc_input, doc_mask, doc_seg)

    @torch.no_grad()
    def predict_step(self, doc_input, doc_mask, doc_seg):
        """Scores each passage and pools over passages (ptBERTMaxP.py:67-96)."""
        _lib.require_cuda(doc_input, doc_mask, doc_seg)
        batch_size = doc_input.shape[0]
        num_passages = self.extractor.config["numpassages"]
        maxseqlen = self.extractor.config["maxseqlen"]

        passage_position = (doc_mask * doc_seg).sum(dim=-1)  # (B, P)
        passage_mask = (passage_position > 5).long()  # (B, P)

        doc_input = doc_input.reshape([batch_size * num_passages, maxseqlen])
        doc_mask = doc_mask.reshape([batch_size * num_passages, maxseqlen])
        doc_seg = doc_seg.reshape([batch_size * num_passages, maxseqlen])

        passage_scores = self.engine().logits(doc_input, doc_mask, doc_seg)[:, 1]
        passage_scores = passage_scores.reshape([batch_size, num_passages])

        if self.config["aggregation"] == "max":
            passage_scores = passage_scores.max(dim=1)[0]  # (batch size, )
        elif self.config["aggregation"] == "first":
            passage_scores = passage_scores[:, 0]
        elif self.config["aggregation"] == "sum":
            passage_scores = torch.sum(passage_mask * passage_scores, dim=1)
        elif self.config["aggregation"] == "avg":
            passage_scores = torch.sum(passage_mask * passage_scores, dim=1) / torch.sum(passage_mask)
        else:
            raise ValueError("Unknown aggregation method: {}".format(self.config["aggregation"]))
        return passage_scores


@Reranker.register
class PTBERTMaxP(Reranker):
    """PyTorch-API BERT-MaxP (monoBERT when numpassages=1).

    Deeper Text Understanding for IR with Contextual Neural Language Modeling. Zhuyun Dai and Jamie Callan. SIGIR 2019."""

    module_name = "ptBERTMaxP"

    dependencies = [
        Dependency(key="extractor", module="extractor", name="bertpassage"),
        Dependency(key="trainer", module="trainer", name="pytorch"),
    ]
    config_spec = [
        ConfigOption("pretrained", "bert-base-uncased",
                     "Pretrained model: bert-base-uncased, bert-base-msmarco, or HuggingFace supported BERT models"),
        ConfigOption("aggregation", "max"),
        ConfigOption("hidden_dropout_prob", 0.1, "The dropout probability of BERT-like model's hidden layers."),
        ConfigOption("precision", "bf16x3", "tensor-core operand mode: bf16x3 (parity, 3 products) or bf16 (fast, ~2e-2 rel. error)"),
    ]

    def build_model(self):
        self.model = PTBERTMaxP_Class(self.extractor, self.config)
        return self.model

    def score(self, d):
        return [
            self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1),
            self.model(d["neg_bert_input"], d["neg_mask"], d["neg_seg"]).view(-1),
        ]

    def test(self, d):
        return self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1)

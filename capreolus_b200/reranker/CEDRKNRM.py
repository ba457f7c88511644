"""CEDR-KNRM behind the reference's module API (``capreolus/reranker/CEDRKNRM.py``).  SURVEY.md §8(f) rank 2.

``CEDRKNRM_Class`` keeps a HF ``BertModel`` as ``self.bert`` exactly like the reference (same ``state_dict`` keys:
``bert.*``, ``kernels.kernels.{i}.mu|sigma``, ``combine.{0,1}``, ``one``, ``zero``).  In eval mode ``forward`` runs the sm_100a
encoder with hidden-state output (``capr_bert_forward_hidden``) and the fused similarity / kernel-pooling / combine head
(``capr_cedrknrm_head``) instead of the torch ops."""
from __future__ import annotations

import math

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption, Dependency
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker.common import RbfKernelBank
from capreolus_b200.reranker.ptBERTMaxP import BertEngine, default_seqs_per_call

_CLS = {None: 0, "avg": 1, "max": 2}


def parse_intlist(v):
    """profane's ``intlist`` value type: a list, ``"1,2,3"`` or ``"start..stop,step"`` (inclusive range, CEDRKNRM.py:203)."""
    if isinstance(v, (list, tuple)):
        return [int(x) for x in v]
    v = str(v)
    if ".." in v:
        rng, _, step = v.partition(",")
        lo, hi = rng.split("..")
        return list(range(int(lo), int(hi) + 1, int(step) if step else 1))
    return [int(x) for x in v.split(",") if x != ""]


class CEDRKNRM_Class(nn.Module):
    """``CEDRKNRM_Class`` (capreolus/reranker/CEDRKNRM.py:14-171)."""

    def __init__(self, extractor, config, *args, **kwargs):
        super().__init__(*args, **kwargs)
        import transformers

        self.extractor = extractor
        self.config = config
        pretrained = config["pretrained"]
        if isinstance(pretrained, dict) and pretrained.get("model_type") == "electra":
            # offline extension: an ElectraConfig dict -> random-init Electra encoder
            cfg_kw = {k: v for k, v in pretrained.items() if k != "model_type"}
            self.bert = transformers.ElectraModel(transformers.ElectraConfig(**{**cfg_kw, "hidden_dropout_prob": config["hidden_dropout_prob"],
                                                                              "output_hidden_states": True}))
        elif isinstance(pretrained, dict):
            # offline extension: a BertConfig dict -> random-init encoder (there is no network for checkpoints here)
            self.bert = transformers.BertModel(transformers.BertConfig(**{**pretrained, "hidden_dropout_prob": config["hidden_dropout_prob"],
                                                                        "output_hidden_states": True}))
        elif isinstance(pretrained, str) and "electra" in pretrained:  # CEDRKNRM.py:20-27 (the reference default)
            name = {"electra-base-msmarco": "Capreolus/electra-base-msmarco", "electra-base": "google/electra-base-discriminator"}.get(pretrained, pretrained)
            self.bert = transformers.ElectraModel.from_pretrained(name, hidden_dropout_prob=config["hidden_dropout_prob"], output_hidden_states=True)
        elif pretrained == "bert-base-msmarco":
            self.bert = transformers.BertModel.from_pretrained("Capreolus/bert-base-msmarco", hidden_dropout_prob=config["hidden_dropout_prob"],
                                                               output_hidden_states=True)
        else:
            self.bert = transformers.BertModel.from_pretrained(pretrained, hidden_dropout_prob=config["hidden_dropout_prob"], output_hidden_states=True)

        self.hidden_size = self.bert.config.hidden_size
        mus = list(self.config["mus"]) + [1.0]
        sigmas = [self.config["sigma"] for _ in self.config["mus"]] + [0.01]
        self.kernels = RbfKernelBank(mus, sigmas, dim=1, requires_grad=self.config["gradkernels"])
        self.simmat_layers = parse_intlist(self.config["simmat_layers"])

        if -1 in self.simmat_layers:
            assert len(self.simmat_layers) == 1
            assert self.config["cls"] is not None
            self._compute_simmat = False
            combine_size = 0
        else:
            self._compute_simmat = True
            combine_size = self.kernels.count() * len(self.simmat_layers)

        assert self.config["cls"] in ("avg", "max", None)
        if self.config["cls"]:
            combine_size += self.hidden_size

        # use weight init from PyTorch 0.4 (CEDRKNRM.py:61-71)
        if config["combine_hidden"] == 0:
            combine_steps = [nn.Linear(combine_size, 1)]
            stdv = 1.0 / math.sqrt(combine_steps[0].weight.size(1))
            combine_steps[0].weight.data.uniform_(-stdv, stdv)
        else:
            combine_steps = [nn.Linear(combine_size, config["combine_hidden"]), nn.Linear(config["combine_hidden"], 1)]
            stdv = 1.0 / math.sqrt(combine_steps[0].weight.size(1))
            combine_steps[0].weight.data.uniform_(-stdv, stdv)
            stdv = 1.0 / math.sqrt(combine_steps[-1].weight.size(1))
            combine_steps[-1].weight.data.uniform_(-stdv, stdv)
        self.combine = nn.Sequential(*combine_steps)

        self.num_passages = extractor.config["numpassages"]
        self.maxseqlen = extractor.config["maxseqlen"]
        self.maxqlen = extractor.config["maxqlen"] + 1  # [SEP] counts as a query position (CEDRKNRM.py:78-80)
        self.maxdoclen = self.maxseqlen - 1
        self.one = nn.Parameter(torch.ones(1), requires_grad=False)
        self.zero = nn.Parameter(torch.zeros(1), requires_grad=False)
        self.precision = config.get("precision", "bf16x3") if hasattr(config, "get") else "bf16x3"
        self.max_seqs_per_call = default_seqs_per_call()
        self._engine, self._engine_key, self._head_ws = None, None, None

    def engine(self) -> BertEngine:
        params = list(self.bert.parameters())
        key = (params[0].device, tuple(p._version for p in params), params[0].data_ptr())
        if self._engine is None or key != self._engine_key:
            self._engine = BertEngine(self.bert, self.precision, max_seqs_per_call=self.max_seqs_per_call)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def _run(self, bert_input, bert_mask, bert_segments, want_feats=False):
        _lib.require_cuda(bert_input, bert_mask, bert_segments)
        lib = _lib.lib()
        B = bert_input.shape[0]
        P, L, H = self.num_passages, self.maxseqlen, self.hidden_size
        ids = bert_input.reshape(B * P, L).long().contiguous()
        mask = bert_mask.reshape(B * P, L).long().contiguous()
        seg = bert_segments.reshape(B * P, L).long().contiguous()
        mu, sigma = self.kernels.stacked()
        K = mu.shape[0]
        layers = self.simmat_layers if self._compute_simmat else []
        cls_mode = _CLS[self.config["cls"]]
        n_enc = self.bert.config.num_hidden_layers
        want = list(layers) + ([n_enc] if cls_mode and n_enc not in layers else [])  # hidden_states[-1] carries the [CLS] feature
        last_slot = want.index(n_enc) if cls_mode else -1
        F = lib.capr_cedrknrm_feature_dim(H, len(layers), K, cls_mode)
        fc1 = self.combine[0]
        hidden = self.config["combine_hidden"]
        fc2 = self.combine[1] if hidden else None
        feats = torch.empty((B, F), dtype=torch.float32, device=ids.device)
        scores = torch.empty((B, 1), dtype=torch.float32, device=ids.device)
        eng = self.engine()
        docs_per_call = max(1, self.max_seqs_per_call // P)
        for lo in range(0, B, docs_per_call):
            nb = min(docs_per_call, B - lo)
            sl = slice(lo * P, (lo + nb) * P)
            hs = eng.hidden_states(ids[sl], mask[sl], seg[sl], want)  # [len(want), nb*P*L, H]
            need = lib.capr_cedrknrm_workspace_bytes(nb * P, self.maxqlen - 1, len(layers), K, hidden)
            if self._head_ws is None or self._head_ws.numel() < need or self._head_ws.device != ids.device:
                self._head_ws = torch.empty(max(need, 256), dtype=torch.uint8, device=ids.device)
            _lib.check(lib.capr_cedrknrm_head(
                hs.data_ptr() if layers else None, len(layers), hs[last_slot].data_ptr() if cls_mode else None, mask[sl].data_ptr(), seg[sl].data_ptr(),
                nb, P, L, H, self.maxqlen - 1, mu.data_ptr(), sigma.data_ptr(), K, cls_mode, fc1.weight.data_ptr(), fc1.bias.data_ptr(), hidden,
                _lib.ptr(fc2.weight if fc2 is not None else None), _lib.ptr(fc2.bias if fc2 is not None else None), feats[lo:lo + nb].data_ptr(),
                scores[lo:lo + nb].data_ptr(), self._head_ws.data_ptr(), self._head_ws.numel(), _lib.current_stream(ids.device)))
        return scores, feats

    def features(self, bert_input, bert_mask, bert_segments):
        """The ``[B, F]`` tensor the reference feeds to ``self.combine`` (CEDRKNRM.py:160-168); tests."""
        return self._run(bert_input, bert_mask, bert_segments)[1]

    def forward(self, bert_input, bert_mask, bert_segments):
        if self.training:
            raise NotImplementedError("capreolus_b200 CEDRKNRM: only inference (model.eval()) is implemented; BERT training is out of scope")
        return self._run(bert_input, bert_mask, bert_segments)[0]


@Reranker.register
class CEDRKNRM(Reranker):
    """CEDR-KNRM (BERT-KNRM when cls=None).  CEDR: Contextualized Embeddings for Document Ranking.
    Sean MacAvaney, Andrew Yates, Arman Cohan, and Nazli Goharian. SIGIR 2019."""

    module_name = "CEDRKNRM"

    dependencies = [
        Dependency(key="extractor", module="extractor", name="pooledbertpassage"),
        Dependency(key="trainer", module="trainer", name="pytorch"),
    ]
    config_spec = [
        ConfigOption("pretrained", "electra-base", "Pretrained model: bert-base-uncased, bert-base-msmarco, electra-base, or electra-base-msmarco"),
        ConfigOption("mus", [-0.9, -0.7, -0.5, -0.3, -0.1, 0.1, 0.3, 0.5, 0.7, 0.9], "mus", value_type="floatlist"),
        ConfigOption("sigma", 0.1, "sigma"),
        ConfigOption("gradkernels", True, "tune mus and sigmas"),
        ConfigOption("hidden_dropout_prob", 0.1, "The dropout probability of BERT-like model's hidden layers."),
        ConfigOption("simmat_layers", "0..12,1", "Layer outputs to include in similarity matrix", value_type="intlist"),
        ConfigOption("combine_hidden", 1024, "Hidden size to use with combination FC layer (0 to disable)"),
        ConfigOption("cls", "avg", "Handling of CLS token: avg, max, or None"),
        ConfigOption("precision", "bf16x3", "tensor-core operand mode of the encoder: bf16x3 (parity) or bf16"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = CEDRKNRM_Class(self.extractor, self.config)
        return self.model

    def score(self, d):
        return [
            self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1),
            self.model(d["neg_bert_input"], d["neg_mask"], d["neg_seg"]).view(-1),
        ]

    def test(self, d):
        return self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1)

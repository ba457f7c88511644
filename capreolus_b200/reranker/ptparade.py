"""PARADE (transformer aggregation) behind the reference's module API (``capreolus/reranker/ptparade.py``).  SURVEY.md §8(f) rank 2.

``PTParade_Class`` keeps the reference's submodules and parameter names (``bert.*`` = HF ``BertModel``, ``transformer_layer_1/2`` =
HF ``BertLayer``, ``linear``, ``initial_cls_embedding``, ``full_position_embeddings``).  In eval mode ``forward`` runs the passage
encoder on the sm_100a engine (``capr_bert_forward_hidden``: last hidden state) and the aggregation head (``capr_parade_head``:
[CLS] gather + position embeddings -> two BertLayers on the same tcgen05 GEMM / attention kernels -> Linear)."""
from __future__ import annotations

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption, Dependency
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker.ptBERTMaxP import BertEngine, default_seqs_per_call


class PTParade_Class(nn.Module):
    """``PTParade_Class`` (capreolus/reranker/ptparade.py:10-78)."""

    def __init__(self, extractor, config, *args, **kwargs):
        super().__init__(*args, **kwargs)
        import transformers
        from transformers.models.bert.modeling_bert import BertLayer

        self.extractor = extractor
        self.config = config
        pretrained = config["pretrained"]
        if isinstance(pretrained, dict):
            # offline extension: a BertConfig dict -> random-init encoder (there is no network for checkpoints here)
            self.bert = transformers.BertModel(transformers.BertConfig(**pretrained))
        elif pretrained == "bert-base-msmarco":
            self.bert = transformers.BertModel.from_pretrained("Capreolus/bert-base-msmarco")
        elif pretrained == "bert-base-uncased":
            self.bert = transformers.BertModel.from_pretrained("bert-base-uncased")
        elif pretrained == "electra-base-msmarco":
            raise ValueError("capreolus_b200 ptparade: electra-base-msmarco is not a BERT encoder (Electra variants are out of scope)")
        else:
            raise ValueError(
                f"unsupported model: {config['pretrained']}; need to ensure correct tokenizers will be used before arbitrary hgf models are supported"
            )

        self.transformer_layer_1 = BertLayer(self.bert.config)
        self.transformer_layer_2 = BertLayer(self.bert.config)
        self.num_passages = extractor.config["numpassages"]
        self.maxseqlen = extractor.config["maxseqlen"]
        self.linear = nn.Linear(self.bert.config.hidden_size, 1)

        if config["aggregation"] in ("max", "avg", "attn"):
            raise NotImplementedError()  # as in the reference (ptparade.py:32-37)
        elif config["aggregation"] == "transformer":
            input_embeddings = self.bert.get_input_embeddings()
            cls_token_id = torch.tensor([[101]])  # hardcoded [CLS] id, as in the reference
            initial_cls = input_embeddings(cls_token_id).view(1, self.bert.config.hidden_size)
            pos = torch.zeros((1, self.num_passages + 1, self.bert.config.hidden_size), requires_grad=True, dtype=torch.float)
            torch.nn.init.normal_(pos, mean=0.0, std=0.02)
            self.initial_cls_embedding = nn.Parameter(initial_cls.detach().clone(), requires_grad=True)
            self.full_position_embeddings = nn.Parameter(pos.detach().clone(), requires_grad=True)
        else:
            raise ValueError(f"unknown aggregation type: {self.config['aggregation']}")
        self.precision = config.get("precision", "bf16x3") if hasattr(config, "get") else "bf16x3"
        self.max_seqs_per_call = default_seqs_per_call()
        self._engine, self._engine_key, self._agg, self._agg_key, self._ws = None, None, None, None, None

    @staticmethod
    def _key(params):
        return (params[0].device, tuple(p._version for p in params), params[0].data_ptr())

    def engine(self) -> BertEngine:
        key = self._key(list(self.bert.parameters()))
        if self._engine is None or key != self._engine_key:
            self._engine, self._engine_key = BertEngine(self.bert, self.precision, max_seqs_per_call=self.max_seqs_per_call), key
        return self._engine

    def aggregator(self) -> BertEngine:
        key = self._key(list(self.transformer_layer_1.parameters()) + list(self.transformer_layer_2.parameters()))
        if self._agg is None or key != self._agg_key:
            self._agg, self._agg_key = BertEngine.from_layers([self.transformer_layer_1, self.transformer_layer_2], self.bert.config, self.precision), key
        return self._agg

    @torch.no_grad()
    def _run(self, doc_input, doc_mask, doc_seg, want_aggregated=False):
        _lib.require_cuda(doc_input, doc_mask, doc_seg)
        lib = _lib.lib()
        B = doc_input.shape[0]
        P, L, H = self.num_passages, self.maxseqlen, self.bert.config.hidden_size
        ids = doc_input.reshape(B * P, L).long().contiguous()
        mask = doc_mask.reshape(B * P, L).long().contiguous()
        seg = doc_seg.reshape(B * P, L).long().contiguous()
        eng, agg = self.engine(), self.aggregator()
        scores = torch.empty((B, 1), dtype=torch.float32, device=ids.device)
        aggregated = torch.empty((B, H), dtype=torch.float32, device=ids.device) if want_aggregated else None
        n_enc = self.bert.config.num_hidden_layers
        docs_per_call = max(1, self.max_seqs_per_call // P)
        cls0 = self.initial_cls_embedding.detach().float().contiguous()
        pos = self.full_position_embeddings.detach().float().contiguous()
        for lo in range(0, B, docs_per_call):
            nb = min(docs_per_call, B - lo)
            sl = slice(lo * P, (lo + nb) * P)
            last = eng.hidden_states(ids[sl], mask[sl], seg[sl], [n_enc])  # [1, nb*P*L, H] = last_hidden_state
            need = lib.capr_parade_workspace_bytes(agg.handle, nb, P)
            if self._ws is None or self._ws.numel() < need or self._ws.device != ids.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=ids.device)
            with torch.cuda.device(ids.device):
                _lib.check(lib.capr_parade_head(agg.handle, last.data_ptr(), nb, P, L, cls0.data_ptr(), pos.data_ptr(), self.linear.weight.data_ptr(),
                                                self.linear.bias.data_ptr(), scores[lo:lo + nb].data_ptr(),
                                                _lib.ptr(aggregated[lo:lo + nb] if aggregated is not None else None), self._ws.data_ptr(),
                                                self._ws.numel(), _lib.current_stream(ids.device)))
        return scores, aggregated

    def aggregate_using_transformer_output(self, doc_input, doc_mask, doc_seg):
        """``transformer_out_2[:, 0, :]`` (ptparade.py:67), ``[B,H]``; tests."""
        return self._run(doc_input, doc_mask, doc_seg, want_aggregated=True)[1]

    def forward(self, doc_input, doc_mask, doc_seg):
        if self.training:
            raise NotImplementedError("capreolus_b200 ptparade: only inference (model.eval()) is implemented; BERT training is out of scope")
        return self._run(doc_input, doc_mask, doc_seg)[0]


@Reranker.register
class PTParade(Reranker):
    """PyTorch implementation of PARADE.

    PARADE: Passage Representation Aggregation for Document Reranking.
    Canjia Li, Andrew Yates, Sean MacAvaney, Ben He, and Yingfei Sun. arXiv 2020."""

    module_name = "ptparade"

    dependencies = [
        Dependency(key="extractor", module="extractor", name="pooledbertpassage"),
        Dependency(key="trainer", module="trainer", name="pytorch"),
    ]
    config_spec = [
        ConfigOption("pretrained", "bert-base-uncased", "Pretrained model: bert-base-uncased, bert-base-msmarco, or electra-base-msmarco"),
        ConfigOption("aggregation", "transformer"),
        ConfigOption("precision", "bf16x3", "tensor-core operand mode of the encoders: bf16x3 (parity) or bf16"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = PTParade_Class(self.extractor, self.config)
        return self.model

    def score(self, d):
        return [
            self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1),
            self.model(d["neg_bert_input"], d["neg_mask"], d["neg_seg"]).view(-1),
        ]

    def test(self, d):
        return self.model(d["pos_bert_input"], d["pos_mask"], d["pos_seg"]).view(-1)

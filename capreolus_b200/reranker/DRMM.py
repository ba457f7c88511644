"""DRMM behind the reference's module API (``capreolus/reranker/DRMM.py``), scored by ``capr_drmm_forward``."""
from __future__ import annotations

import torch
import torch.nn as nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker import common
from capreolus_b200.reranker.common import SimilarityMatrix, _ids, create_emb_layer

_HIST = {"CH": 0, "NH": 1, "LCH": 2}
_GATE = {"IDF": 0, "TV": 1}


class DRMM_class(nn.Module):
    """``DRMM_class`` (capreolus/reranker/DRMM.py:13-116); parameters ``ffw.{0,2}``, ``gates``, ``output_layer``."""

    def __init__(self, extractor, config):
        super(DRMM_class, self).__init__()
        self.nbins = config["nbins"]
        self.nodes = config["nodes"]
        self.hist_type = config["histType"]
        self.gate_type = config["gateType"]

        self.embedding = create_emb_layer(extractor.embeddings, non_trainable=True)
        self.simmat = SimilarityMatrix(self.embedding)
        self._prepared = self.simmat._prepared

        self.ffw = nn.Sequential(nn.Linear(self.nbins + 1, self.nodes), nn.Tanh(), nn.Linear(self.nodes, 1), nn.Tanh())
        emb_dim = self.embedding.weight.size(-1)
        if self.gate_type == "IDF":
            self.gates = nn.Linear(1, 1, bias=False)
        elif self.gate_type == "TV":
            self.gates = nn.Linear(emb_dim, 1, bias=False)
        else:
            raise ValueError("Invalid value for gateType: gateType should be either IDF or TV")
        if self.hist_type not in _HIST:
            raise ValueError("Invalid value for histType: histType should be 'CH', 'NH', or 'LCH'")
        self.output_layer = nn.Linear(1, 1)

        # initialize FC and gate weight in the same way as MatchZoo (DRMM.py:36-39)
        nn.init.uniform_(self.ffw[0].weight, -0.1, 0.1)
        nn.init.uniform_(self.ffw[2].weight, -0.1, 0.1)
        nn.init.uniform_(self.gates.weight, -0.01, 0.01)
        # the exact fp32 bin bounds the reference compares against (DRMM.py:63)
        self.register_buffer("_nosave_bin_ub", torch.linspace(-1, 1, self.nbins + 1)[1:].contiguous(), persistent=False)

    def _run(self, sentence, query_sentence, query_idf, want_hist=False):
        _lib.require_cuda(sentence, query_sentence, query_idf)
        if self.training and torch.is_grad_enabled() and not want_hist and any(p.requires_grad for p in self.parameters()):  # eval mode scores with the inference kernels whatever the grad mode
            # training: histogram from the CUDA engine (frozen embedding, no gradient needed), parameterised tail in torch (train_heads.py)
            from capreolus_b200.reranker import train_heads

            with torch.no_grad():
                hist = self._run(sentence, query_sentence, query_idf, want_hist=True)[1]
            return train_heads.drmm_forward(self, hist, _ids(query_sentence), query_idf), None
        q, d = _ids(query_sentence), _ids(sentence)
        B, Q = q.shape
        D = d.shape[1]
        idf = query_idf.float().contiguous()
        raw = self.embedding.weight.detach().contiguous()
        scores = torch.empty((B, 1), dtype=torch.float32, device=q.device)
        hist = torch.empty((B, Q, self.nbins + 1), dtype=torch.float32, device=q.device) if want_hist else None
        if common.use_tensor_cores(D, raw.shape[1]) and self.nbins <= 31:
            hi, lo = self._prepared.get_bf16()
            _lib.check(_lib.lib().capr_drmm_forward_tc(
                q.data_ptr(), d.data_ptr(), idf.data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], hi.shape[1], raw.data_ptr(),
                raw.shape[1], self.nbins, self._nosave_bin_ub.data_ptr(), _HIST[self.hist_type], _GATE[self.gate_type],
                self.ffw[0].weight.data_ptr(), self.ffw[0].bias.data_ptr(), self.nodes, self.ffw[2].weight.data_ptr(),
                self.ffw[2].bias.data_ptr(), self.gates.weight.data_ptr(), self.output_layer.weight.data_ptr(),
                self.output_layer.bias.data_ptr(), scores.data_ptr(), _lib.ptr(hist), _lib.current_stream(q.device)))
            return scores, hist
        table = self._prepared.get()
        _lib.check(_lib.lib().capr_drmm_forward(
            q.data_ptr(), d.data_ptr(), idf.data_ptr(), B, Q, D, table.data_ptr(), table.shape[0], table.shape[1], raw.data_ptr(),
            raw.shape[1], self.nbins, self._nosave_bin_ub.data_ptr(), _HIST[self.hist_type], _GATE[self.gate_type],
            self.ffw[0].weight.data_ptr(), self.ffw[0].bias.data_ptr(), self.nodes, self.ffw[2].weight.data_ptr(),
            self.ffw[2].bias.data_ptr(), self.gates.weight.data_ptr(), self.output_layer.weight.data_ptr(),
            self.output_layer.bias.data_ptr(), scores.data_ptr(), _lib.ptr(hist), _lib.current_stream(q.device)))
        return scores, hist

    def _hist_map(self, queries, documents, d_masks=None):
        """The transformed matching histogram ``[B,Q,nbins+1]`` (DRMM.py:41-81); takes token ids like the reference."""
        idf = torch.zeros(queries.shape, dtype=torch.float32, device=queries.device)
        return self._run(documents, queries, idf, want_hist=True)[1]

    def forward(self, sentence, query_sentence, query_idf):
        return self._run(sentence, query_sentence, query_idf)[0]


@Reranker.register
class DRMM(Reranker):
    """Jiafeng Guo, Yixing Fan, Qingyao Ai, and W. Bruce Croft. 2016. A Deep Relevance Matching Model for Ad-hoc Retrieval. In CIKM'16."""

    module_name = "DRMM"

    config_spec = [
        ConfigOption("nbins", 29, "number of bins in matching histogram"),
        ConfigOption("nodes", 5, "hidden layer dimension for feed forward matching network"),
        ConfigOption("histType", "LCH", "histogram type: CH, NH, or LCH"),
        ConfigOption("gateType", "IDF", "term gate type: TV or IDF"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = DRMM_class(self.extractor, self.config)
        return self.model

    def score(self, d):
        query_idf = d["query_idf"]
        query_sentence = d["query"]
        pos_sentence, neg_sentence = d["posdoc"], d["negdoc"]
        return [
            self.model(pos_sentence, query_sentence, query_idf).view(-1),
            self.model(neg_sentence, query_sentence, query_idf).view(-1),
        ]

    def test(self, d):
        query_idf = d["query_idf"]
        query_sentence = d["query"]
        pos_sentence = d["posdoc"]
        return self.model(pos_sentence, query_sentence, query_idf).view(-1)

"""DRMMTKS behind the reference's module API (``capreolus/reranker/DRMMTKS.py``), scored by ``capr_drmmtks_forward_tc``.

SURVEY.md §8(f) rank 1.  Parameter names follow the reference (``ffw.0``, ``gates``, ``output_layer``,
``embedding.weight``), so checkpoints interchange."""
from __future__ import annotations

import torch
import torch.nn as nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker.common import SimilarityMatrix, _ids, create_emb_layer


class DRMMTKS_class(nn.Module):
    """``DRMMTKS_class`` (capreolus/reranker/DRMMTKS.py:13-63)."""

    def __init__(self, extractor, config):
        super(DRMMTKS_class, self).__init__()
        self.topk = config["topk"]
        self.gate_type = config["gateType"]

        self.embedding = create_emb_layer(extractor.embeddings, non_trainable=config["freezeemb"])
        self.simmat = SimilarityMatrix(self.embedding)
        self._prepared = self.simmat._prepared

        self.ffw = nn.Sequential(nn.Linear(self.topk, 1), nn.Tanh())
        gate_inp_dim = 1 if self.gate_type == "IDF" else self.embedding.weight.size(-1)
        self.gates = nn.Linear(gate_inp_dim, 1, bias=False)
        self.output_layer = nn.Linear(1, 1)

        # initialize FC and gate weight in the same way as MatchZoo (DRMMTKS.py:28-30)
        nn.init.uniform_(self.ffw[0].weight, -0.1, 0.1)
        nn.init.uniform_(self.gates.weight, -0.01, 0.01)

    def _run(self, doc, query, query_idf, want_topk=False):
        _lib.require_cuda(doc, query, query_idf)
        if self.gate_type != "IDF":
            # DRMMTKS.py:59 hands the int64 token ids to _term_gate, whose TV branch applies Linear(E,1) to them (l.43):
            # the reference raises there, so there is no behaviour to reproduce
            raise ValueError("Invalid value for gateType: DRMMTKS supports gateType='IDF' only (the reference's 'TV' branch cannot run)")
        if self.training and torch.is_grad_enabled() and not want_topk and any(p.requires_grad for p in self.parameters()):  # eval mode scores with the inference kernels whatever the grad mode
            # training: top-k cosines from the CUDA engine, parameterised tail in torch (train_heads.py)
            from capreolus_b200.reranker import train_heads

            if self.embedding.weight.requires_grad:
                raise NotImplementedError("capreolus_b200 DRMMTKS: freezeemb=False (gradient to the embedding table) is not implemented")
            with torch.no_grad():
                topk = self._run(doc, query, query_idf, want_topk=True)[1]
            return train_heads.drmmtks_forward(self, topk, _ids(query), query_idf), None
        q, d = _ids(query), _ids(doc)
        B, Q = q.shape
        D = d.shape[1]
        idf = query_idf.float().contiguous()
        hi, lo = self._prepared.get_bf16()
        scores = torch.empty((B, 1), dtype=torch.float32, device=q.device)
        topk = torch.empty((B, Q, self.topk), dtype=torch.float32, device=q.device) if want_topk else None
        _lib.check(_lib.lib().capr_drmmtks_forward_tc(
            q.data_ptr(), d.data_ptr(), idf.data_ptr(), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], self.embedding.weight.shape[1],
            hi.shape[1], self.topk, self.ffw[0].weight.data_ptr(), self.ffw[0].bias.data_ptr(), self.gates.weight.data_ptr(),
            self.output_layer.weight.data_ptr(), self.output_layer.bias.data_ptr(), scores.data_ptr(), _lib.ptr(topk),
            _lib.current_stream(q.device)))
        return scores, topk

    def topk_similarities(self, doc, query):
        """``torch.topk(cos_mat, k)[0]`` of DRMMTKS.py:55-56, ``[B,Q,k]`` (tests)."""
        idf = torch.zeros(query.shape, dtype=torch.float32, device=query.device)
        return self._run(doc, query, idf, want_topk=True)[1]

    def forward(self, doc, query, query_idf):
        return self._run(doc, query, query_idf)[0]


@Reranker.register
class DRMMTKS(Reranker):
    """Jiafeng Guo, Yixing Fan, Qingyao Ai, and W. Bruce Croft. 2016. A Deep Relevance Matching Model for Ad-hoc Retrieval. In CIKM'16."""

    # reference: https://github.com/NTMC-Community/MatchZoo-py/blob/master/matchzoo/models/drmmtks.py
    module_name = "DRMMTKS"

    config_spec = [
        ConfigOption("topk", 10, "number of bins in matching histogram"),
        ConfigOption("gateType", "IDF", "term gate type: TV or IDF"),
        ConfigOption("freezeemb", True, "term gate type: TV or IDF"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = DRMMTKS_class(self.extractor, self.config)
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

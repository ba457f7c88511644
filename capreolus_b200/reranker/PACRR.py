"""PACRR behind the reference's module API (``capreolus/reranker/PACRR.py``), scored by ``capr_pacrr_forward``."""
from __future__ import annotations

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker import common
from capreolus_b200.reranker.common import SimilarityMatrix, _ids, create_emb_layer, device_pointer_array

_NONLIN = {"none": 0, "relu": 1, "tanh": 2}


class PACRRConvMax2dModule(torch.nn.Module):
    """Parameter holder for one n-gram module (``conv.weight [F,1,n,n]``, ``conv.bias [F]``; PACRR.py:57-82)."""

    def __init__(self, shape, n_filters, k, channels):
        super().__init__()
        self.shape = shape
        self.conv = torch.nn.Conv2d(channels, n_filters, shape)
        self.k = k
        self.channels = channels


class PACRR_class(nn.Module):
    """``PACRR_class`` (capreolus/reranker/PACRR.py:13-54); parameters ``ngrams.{i}.conv``, ``linear{1,2,3}``."""

    def __init__(self, extractor, config):
        super(PACRR_class, self).__init__()
        p = config
        self.p = p
        self.extractor = extractor
        self.embedding_dim = extractor.embeddings.shape[1]
        self.embedding = create_emb_layer(extractor.embeddings, non_trainable=True)
        self.simmat = SimilarityMatrix(self.embedding)
        self._prepared = self.simmat._prepared

        self.ngrams = nn.ModuleList()
        for ng in range(p["mingram"], p["maxgram"] + 1):
            self.ngrams.append(PACRRConvMax2dModule(ng, p["nfilters"], k=p["kmax"], channels=1))

        qterm_size = len(self.ngrams) * p["kmax"] + (1 if p["idf"] else 0)
        self.linear1 = torch.nn.Linear(extractor.config["maxqlen"] * qterm_size, p["combine"])
        self.linear2 = torch.nn.Linear(p["combine"], p["combine"])
        self.linear3 = torch.nn.Linear(p["combine"], 1)
        if p["nonlinearity"] not in _NONLIN:
            raise ValueError("nonlinearity must be one of: none, relu, tanh")
        nonlinearity = {"none": torch.nn.Identity, "relu": torch.nn.ReLU, "tanh": torch.nn.Tanh}[p["nonlinearity"]]
        # same container (and therefore the same duplicate ``combine.{0,2,4}`` state_dict keys) as PACRR.py:40
        self.combine = torch.nn.Sequential(self.linear1, nonlinearity(), self.linear2, nonlinearity(), self.linear3)

    def _run(self, sentence, query_sentence, query_idf, want_topk=False):
        _lib.require_cuda(sentence, query_sentence)
        if self.training and torch.is_grad_enabled() and not want_topk and any(p.requires_grad for p in self.parameters()):  # eval mode scores with the inference kernels whatever the grad mode
            # training: cosine matrix from the CUDA engine (capr_simmat_forward), Conv2d / max / top-k / MLP in torch (train_heads.py)
            from capreolus_b200.reranker import train_heads

            with torch.no_grad():
                sim = self.simmat(query_sentence, sentence)
            return train_heads.pacrr_forward(self, sim, query_idf), None
        p = self.p
        q, d = _ids(query_sentence), _ids(sentence)
        B, Q = q.shape
        D = d.shape[1]
        if Q != self.extractor.config["maxqlen"]:
            raise ValueError(f"query length {Q} != extractor maxqlen {self.extractor.config['maxqlen']} (PACRR.py:30,50)")
        idf = query_idf.float().contiguous() if p["idf"] else None
        ws = [ng.conv.weight.detach().contiguous() for ng in self.ngrams]
        bs = [ng.conv.bias.detach().contiguous() for ng in self.ngrams]
        scores = torch.empty((B, 1), dtype=torch.float32, device=q.device)
        topk = torch.empty((B, Q, len(self.ngrams) * p["kmax"]), dtype=torch.float32, device=q.device) if want_topk else None
        if common.use_tensor_cores(D, self.embedding_dim, max_doclen=512) and len(self.ngrams) * p["kmax"] <= 12:
            hi, lo = self._prepared.get_bf16()
            _lib.check(_lib.lib().capr_pacrr_forward_tc(
                q.data_ptr(), d.data_ptr(), _lib.ptr(idf), B, Q, D, hi.data_ptr(), lo.data_ptr(), hi.shape[0], self.embedding_dim, hi.shape[1],
                p["mingram"], p["maxgram"], p["nfilters"], p["kmax"], device_pointer_array(ws), device_pointer_array(bs),
                self.linear1.weight.data_ptr(), self.linear1.bias.data_ptr(), self.linear2.weight.data_ptr(), self.linear2.bias.data_ptr(),
                self.linear3.weight.data_ptr(), self.linear3.bias.data_ptr(), p["combine"], _NONLIN[p["nonlinearity"]], scores.data_ptr(),
                _lib.ptr(topk), _lib.current_stream(q.device)))
            return scores, topk
        table = self._prepared.get()
        _lib.check(_lib.lib().capr_pacrr_forward(
            q.data_ptr(), d.data_ptr(), _lib.ptr(idf), B, Q, D, table.data_ptr(), table.shape[0], table.shape[1], p["mingram"], p["maxgram"],
            p["nfilters"], p["kmax"], device_pointer_array(ws), device_pointer_array(bs), self.linear1.weight.data_ptr(),
            self.linear1.bias.data_ptr(), self.linear2.weight.data_ptr(), self.linear2.bias.data_ptr(), self.linear3.weight.data_ptr(),
            self.linear3.bias.data_ptr(), p["combine"], _NONLIN[p["nonlinearity"]], scores.data_ptr(), _lib.ptr(topk),
            _lib.current_stream(q.device)))
        return scores, topk

    def ngram_topk(self, sentence, query_sentence):
        """``cat([ng(simmat) for ng in ngrams], dim=2)`` -> ``[B,Q,ngrams*kmax]`` (PACRR.py:46,82)."""
        idf = torch.zeros(query_sentence.shape, dtype=torch.float32, device=query_sentence.device)
        return self._run(sentence, query_sentence, idf, want_topk=True)[1]

    def forward(self, sentence, query_sentence, query_idf):
        return self._run(sentence, query_sentence, query_idf)[0]


@Reranker.register
class PACRR(Reranker):
    """Kai Hui, Andrew Yates, Klaus Berberich, and Gerard de Melo. 2017. PACRR: A Position-Aware Neural IR Model for Relevance Matching. EMNLP 2017."""

    module_name = "PACRR"

    config_spec = [
        ConfigOption("mingram", 1, "minimum length of ngram used"),
        ConfigOption("maxgram", 3, "maximum length of ngram used"),
        ConfigOption("nfilters", 32, "number of filters in convolution layer"),
        ConfigOption("idf", True, "concatenate idf signals to combine relevance score from individual query terms"),
        ConfigOption("kmax", 2, "value of kmax pooling used"),
        ConfigOption("combine", 32, "size of combination layers"),
        ConfigOption("nonlinearity", "relu", "nonlinearity in combination layer: none, relu, or tanh"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = PACRR_class(self.extractor, self.config)
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

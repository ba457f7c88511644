"""ConvKNRM behind the reference's module API (``capreolus/reranker/ConvKNRM.py``), scored by ``capr_convknrm_forward``.

SURVEY.md §8(f) rank 1.  Submodule / parameter names follow the reference (``embeddings.weight``,
``kernels.kernels.{i}.mu|sigma``, ``convs.{n}.0.weight|bias``, ``combine.{0,2}``), so checkpoints interchange.
The Conv1d encoders are folded into projected tables (``capr_convknrm_project``) that are rebuilt lazily when the
conv or embedding weights change."""
from __future__ import annotations

import os

import torch
from torch import nn

from capreolus_b200 import _lib
from capreolus_b200.module import ConfigOption, Dependency
from capreolus_b200.reranker import Reranker
from capreolus_b200.reranker import common
from capreolus_b200.reranker.common import RbfKernelBank, _ids, create_emb_layer, device_pointer_array

SCORETANH = 1  # CAPR_KNRM_SCORETANH
#: pairs whose n-gram representations are staged in HBM at a time (835 KB per pair at the default config)
CHUNK = int(os.environ.get("CAPR_CONVKNRM_CHUNK", "4096"))


class StackedSimilarityMatrix(nn.Module):
    """Holder for API parity with ``StackedSimilarityMatrix`` (capreolus/reranker/common.py:187-221): the cosine views are
    produced and consumed inside ``capr_convknrm_forward`` and never materialised."""

    def __init__(self, padding=0):
        super().__init__()
        self.padding = padding


class ConvKNRM_class(nn.Module):
    """``ConvKNRM_class`` (capreolus/reranker/ConvKNRM.py:13-77)."""

    def __init__(self, extractor, config):
        super(ConvKNRM_class, self).__init__()
        self.p = config
        self.simmat = StackedSimilarityMatrix(padding=getattr(extractor, "pad", 0))
        if self.simmat.padding != 0:
            raise NotImplementedError("capreolus_b200 ConvKNRM: the <pad> token id must be 0 (EmbedText.pad)")
        self.embeddings = create_emb_layer(extractor.embeddings, non_trainable=True)

        mus = [-0.9, -0.7, -0.5, -0.3, -0.1, 0.1, 0.3, 0.5, 0.7, 0.9, 1.0]
        sigmas = [0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.001]
        self.kernels = RbfKernelBank(mus, sigmas, dim=1, requires_grad=config["gradkernels"])

        self.padding, self.convs = nn.ModuleList(), nn.ModuleList()
        for conv_size in range(1, config["maxngram"] + 1):
            if conv_size > 1:
                self.padding.append(nn.ConstantPad1d((0, conv_size - 1), 0))
            else:
                self.padding.append(nn.Sequential())  # identity
            self.convs.append(nn.ModuleList())
            for _ in range(1):
                self.convs[-1].append(nn.Conv1d(self.embeddings.weight.shape[1], config["filters"], conv_size))

        channels = config["maxngram"] ** 2 if config["crossmatch"] else config["maxngram"]
        if config["singlefc"]:
            combine_steps = [nn.Linear(self.kernels.count() * channels, 1)]
        else:
            combine_steps = [nn.Linear(self.kernels.count() * channels, 30), nn.Tanh(), nn.Linear(30, 1)]
        if config["scoretanh"]:
            combine_steps.append(nn.Tanh())
        self.combine = nn.Sequential(*combine_steps)
        self._proj, self._proj_key, self._ws = None, None, None

    def projected_table(self) -> torch.Tensor:
        """``[V, S*F]`` fp32: emb . W_n[:, :, u]^T for every (n, u) (``capr_convknrm_project``); derived data, never saved."""
        emb = self.embeddings.weight
        ws = [c[0].weight for c in self.convs]
        key = tuple((t.data_ptr(), t._version, t.device) for t in [emb] + ws)
        if key != self._proj_key:
            V, E = emb.shape
            F, n = self.p["filters"], self.p["maxngram"]
            proj = torch.empty((V, _lib.lib().capr_convknrm_proj_cols(n, F)), dtype=torch.float32, device=emb.device)
            wc = [w.detach().contiguous() for w in ws]
            _lib.check(_lib.lib().capr_convknrm_project(emb.detach().contiguous().data_ptr(), V, E, device_pointer_array(wc), n, F, proj.data_ptr(),
                                                        _lib.current_stream(emb.device)))
            self._proj, self._proj_key = proj, key
        return self._proj

    def _workspace(self, B, Q, D, K, device):
        chunk = max(1, min(B, CHUNK))
        need = _lib.lib().capr_convknrm_workspace_bytes(chunk, Q, D, self.p["maxngram"], self.p["filters"], K, int(bool(self.p["crossmatch"])))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _run(self, sentence, query_sentence, want_feats=False):
        _lib.require_cuda(sentence, query_sentence)
        if self.training and torch.is_grad_enabled() and not want_feats and any(p.requires_grad for p in self.parameters()):  # eval mode scores with the inference kernels whatever the grad mode
            # training: the Conv1d encoders sit upstream of the cosine, so the whole forward is the torch restatement (train_heads.py)
            from capreolus_b200.reranker import train_heads

            return train_heads.convknrm_forward(self, sentence, query_sentence), None
        q, d = _ids(query_sentence), _ids(sentence)
        B, Q = q.shape
        D = d.shape[1]
        mu, sigma = self.kernels.stacked()
        K = mu.shape[0]
        views = self.p["maxngram"] ** 2 if self.p["crossmatch"] else self.p["maxngram"]
        fc1 = self.combine[0]
        hidden = 0 if self.p["singlefc"] else fc1.out_features
        fc2 = None if self.p["singlefc"] else self.combine[2]
        proj = self.projected_table()
        ws = self._workspace(B, Q, D, K, q.device)
        biases = [c[0].bias.detach().contiguous() for c in self.convs]
        scores = torch.empty((B, 1), dtype=torch.float32, device=q.device)
        feats = torch.empty((B, K * views), dtype=torch.float32, device=q.device) if want_feats else None
        _lib.check(_lib.lib().capr_convknrm_forward(
            q.data_ptr(), d.data_ptr(), B, Q, D, proj.data_ptr(), proj.shape[0], self.p["maxngram"], self.p["filters"], device_pointer_array(biases),
            int(bool(self.p["crossmatch"])), mu.data_ptr(), sigma.data_ptr(), K, fc1.weight.data_ptr(), fc1.bias.data_ptr(), hidden,
            _lib.ptr(fc2.weight if fc2 is not None else None), _lib.ptr(fc2.bias if fc2 is not None else None),
            (SCORETANH if self.p["scoretanh"] else 0) | common.DEBUG_FLAGS, scores.data_ptr(), _lib.ptr(feats), ws.data_ptr(), ws.numel(),
            _lib.current_stream(q.device)))
        return scores, feats

    def kernel_features(self, sentence, query_sentence):
        """The ``[B, K*VIEWS]`` tensor the reference feeds to ``self.combine`` (ConvKNRM.py:75-76); tests."""
        return self._run(sentence, query_sentence, want_feats=True)[1]

    def forward(self, sentence, query_sentence, query_idf):
        return self._run(sentence, query_sentence)[0]


@Reranker.register
class ConvKNRM(Reranker):
    """Zhuyun Dai, Chenyan Xiong, Jamie Callan, and Zhiyuan Liu. 2018. Convolutional Neural Networks for Soft-Matching N-Grams in Ad-hoc Search. In WSDM'18."""

    module_name = "ConvKNRM"

    dependencies = [
        Dependency(key="extractor", module="extractor", name="slowembedtext"),
        Dependency(key="trainer", module="trainer", name="pytorch"),
    ]
    config_spec = [
        ConfigOption("gradkernels", True, "backprop through mus and sigmas"),
        ConfigOption("maxngram", 3, "maximum ngram length considered"),
        ConfigOption("crossmatch", True, "match query and document ngrams of different lengths (e.g., bigram vs. unigram)"),
        ConfigOption("filters", 128, "number of filters used in convolutional layers"),
        ConfigOption("scoretanh", False, "use a tanh on the prediction as in paper (True) or do not use a nonlinearity (False)"),
        ConfigOption("singlefc", True, "use single fully connected layer as in paper (True) or 2 fully connected layers (False)"),
    ]

    def build_model(self):
        if not hasattr(self, "model"):
            self.model = ConvKNRM_class(self.extractor, self.config)
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

    def zero_grad(self, *args, **kwargs):
        self.model.zero_grad(*args, **kwargs)

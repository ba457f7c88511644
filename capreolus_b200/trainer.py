"""The caller of the hot path during training, restated so the pairwise-hinge loop can run where the reference
package is not importable.

``single_train_iteration`` follows ``PytorchTrainer.single_train_iteration`` (``capreolus/trainer/pytorch.py:76-122``)
step for step: ``reranker.score(batch)`` -> loss -> ``backward`` -> every ``gradacc`` batches ``optimizer.step()`` /
``zero_grad()`` -> stop after ``itersize // batch`` batches -> mean of the batch losses.  It is glue: the forward,
the d/dmu, d/dsigma statistics and the hinge loss are the CUDA kernels behind ``reranker.score`` / ``pair_hinge_loss``.
"""
from __future__ import annotations

import os

import torch

from capreolus_b200.reranker.common import pair_hinge_loss, pair_softmax_loss


class PairwiseTrainer:
    def __init__(self, batch=32, itersize=512, gradacc=1, lr=0.001, softmaxloss=False, device="cuda:0"):
        if batch < 1:
            raise ValueError("batch must be >= 1")
        if itersize < batch:
            raise ValueError("itersize must be >= batch")
        if gradacc < 1 or not float(gradacc).is_integer():
            raise ValueError("gradacc must be an integer >= 1")
        if lr <= 0:
            raise ValueError("lr must be > 0")
        self.config = dict(batch=batch, itersize=itersize, gradacc=gradacc, lr=lr, softmaxloss=softmaxloss)
        self.device = torch.device(device)
        self.loss = pair_softmax_loss if softmaxloss else pair_hinge_loss  # trainer/pytorch.py:220-223
        self.optimizer = None

    @property
    def n_batch_per_iter(self):
        return (self.config["itersize"] // self.config["batch"]) or 1  # trainer/__init__.py:74-76

    def prepare(self, reranker):
        model = reranker.model.to(self.device)
        model.train()
        params = [p for p in model.parameters() if p.requires_grad]
        # pytorch.py:205 builds torch.optim.Adam(params, lr); `fused=True` is the same update rule as ONE kernel over all parameters instead
        # of a foreach group per operation (the KNRM iteration is launch-bound: 27 small parameters); CAPR_TRAINER_FUSED_ADAM=0 restores the default
        fused = os.environ.get("CAPR_TRAINER_FUSED_ADAM", "1") != "0" and all(p.is_cuda for p in params)
        self.optimizer = torch.optim.Adam(params, lr=self.config["lr"], **({"fused": True} if fused else {}))
        return model

    def single_train_iteration(self, reranker, train_dataloader, cur_iter=0):
        iter_loss = []
        batches_since_update = 0
        for bi, batch in enumerate(train_dataloader):
            batch = {k: v.to(self.device) if not isinstance(v, list) else v for k, v in batch.items()}
            doc_scores = reranker.score(batch)
            loss = self.loss(doc_scores)
            iter_loss.append(loss)
            loss.backward()
            batches_since_update += 1
            if batches_since_update == self.config["gradacc"]:
                batches_since_update = 0
                self.optimizer.step()
                self.optimizer.zero_grad()
            if (bi + 1) % self.n_batch_per_iter == 0:
                break
        return torch.stack(iter_loss).mean()

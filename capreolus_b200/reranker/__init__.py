"""Drop-in counterpart of ``capreolus/reranker/__init__.py:7-55``: the ``Reranker`` module base class.

A Reranker provides ``build_model()`` (sets ``self.model``), ``score(batch) -> [pos(B,), neg(B,)]`` and
``test(batch) -> (B,)``; the trainer calls those (``trainer/pytorch.py:97,344``).  ``save_weights`` /
``load_weights`` keep the reference's checkpoint format: a pickle of ``state_dict`` without
``embedding.weight`` and ``_nosave_`` entries, plus ``<file>.optimizer``.
"""
from __future__ import annotations

import os
import pickle
from pathlib import Path

from capreolus_b200.module import ConfigOption, Dependency, ModuleBase


class Reranker(ModuleBase):
    module_type = "reranker"
    dependencies = [
        Dependency(key="extractor", module="extractor", name="embedtext"),
        Dependency(key="trainer", module="trainer", name="pytorch"),
    ]

    def add_summary(self, summary_writer, niter):
        for name, weight in self.model.named_parameters():
            summary_writer.add_histogram(name, weight.data.cpu(), niter)

    @staticmethod
    def _saved(key: str) -> bool:
        return "embedding.weight" not in key and "_nosave_" not in key

    def save_weights(self, weights_fn, optimizer):
        weights_fn = Path(weights_fn)
        os.makedirs(weights_fn.parent, exist_ok=True)
        d = {k: v for k, v in self.model.state_dict().items() if self._saved(k)}
        with open(weights_fn, "wb") as outf:
            pickle.dump(d, outf, protocol=-1)
        with open(weights_fn.as_posix() + ".optimizer", "wb") as outf:
            pickle.dump(optimizer.state_dict(), outf, protocol=-1)

    def load_weights(self, weights_fn, optimizer):
        weights_fn = Path(weights_fn)
        with open(weights_fn, "rb") as f:
            d = pickle.load(f)
        cur_keys = set(k for k in self.model.state_dict().keys() if self._saved(k))
        missing = cur_keys - set(d.keys())
        if len(missing) > 0:
            raise RuntimeError("loading state_dict with keys that do not match current model: %s" % missing)
        self.model.load_state_dict(d, strict=False)
        with open(weights_fn.as_posix() + ".optimizer", "rb") as f:
            optimizer.load_state_dict(pickle.load(f))


from capreolus_b200.reranker.KNRM import KNRM, KNRM_class  # noqa: E402,F401
from capreolus_b200.reranker.DRMM import DRMM, DRMM_class  # noqa: E402,F401
from capreolus_b200.reranker.PACRR import PACRR, PACRR_class  # noqa: E402,F401
from capreolus_b200.reranker.ptBERTMaxP import PTBERTMaxP, PTBERTMaxP_Class  # noqa: E402,F401
from capreolus_b200.reranker.DRMMTKS import DRMMTKS, DRMMTKS_class  # noqa: E402,F401
from capreolus_b200.reranker.ConvKNRM import ConvKNRM, ConvKNRM_class  # noqa: E402,F401
from capreolus_b200.reranker.CEDRKNRM import CEDRKNRM, CEDRKNRM_Class  # noqa: E402,F401
from capreolus_b200.reranker.ptparade import PTParade, PTParade_Class  # noqa: E402,F401

__all__ = ["Reranker", "ConfigOption", "Dependency", "KNRM", "KNRM_class", "DRMM", "DRMM_class", "PACRR", "PACRR_class", "PTBERTMaxP", "PTBERTMaxP_Class", "DRMMTKS", "DRMMTKS_class", "ConvKNRM", "ConvKNRM_class", "CEDRKNRM", "CEDRKNRM_Class", "PTParade", "PTParade_Class"]

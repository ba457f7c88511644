import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)


def load_golden(name):
    """npz -> dict; int32 id arrays are widened back to the reference's int64."""
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype == np.int32 and k in ("query", "posdoc", "negdoc", "pos_bert_input", "pos_mask", "pos_seg"):
            v = v.astype(np.int64)
        out[k] = v
    return out


def golden_state(g, variant):
    import torch

    pre = f"{variant}/state/"
    return {k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}


def golden_table(g):
    """Re-derive the embedding table from its seed and check it against the stored checksum."""
    from capreolus_b200 import synthetic

    B, Q, D, V, E = (int(x) for x in g["shape"])
    table = synthetic.embedding_table(V, E, seed=int(g["table_seed"]))
    chk = np.array([table.astype(np.float64).sum(), np.abs(table.astype(np.float64)).sum(), float(table[-1, -1])])
    np.testing.assert_allclose(chk, g["table_checksum"], rtol=1e-12)
    return table


def rel_err(got, want, floor=1e-3):
    """max |got-want| / max(|want|, floor): the 1e-3 'fp32 relative' bar of BASELINE.json, with a floor so
    that scores that happen to be ~0 do not blow the ratio up."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))

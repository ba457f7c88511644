"""GPU: the host-side predict pipeline (H2D / score / D2H overlap) returns exactly what reranker.test returns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Extractor:
    def __init__(self, table, Q, D):
        self.embeddings = table
        self.config = {"maxqlen": Q, "maxdoclen": D}


@pytest.mark.parametrize("n,chunk", [(1000, 256), (513, 512), (7, 64), (2048, 2048)])
def test_pipelined_predict_equals_direct_test(n, chunk):
    from capreolus_b200 import reranker as R, synthetic
    from capreolus_b200.predict import PinnedBatch, PipelinedPredictor

    Q, D, V, E = 32, 512, 3000, 300
    rr = R.KNRM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    rr.build_model().to(DEV).eval()
    host = {k: torch.from_numpy(v) for k, v in synthetic.throughput_batch(n, Q, D, V, seed=9).items()}
    pb = PinnedBatch(host)
    assert pb.n == n and pb.bytes_per_item == (Q + D) * 8 + Q * 4
    pred = PipelinedPredictor(rr, DEV, chunk=chunk)
    with torch.no_grad():
        direct = rr.test({k: v.to(DEV) for k, v in host.items()}).cpu()
    for _ in range(2):  # buffers are reused across calls
        out = pred.predict(pb)
        assert out.shape == (n,) and out.is_pinned()
        assert torch.equal(out.cpu(), direct)


def test_table_is_rebuilt_when_the_embedding_changes():
    """PreparedTable caches derived data keyed on the weight's version counter (finetune=True steps, load_state_dict)."""
    from capreolus_b200 import reranker as R, synthetic

    Q, D, V, E = 8, 40, 500, 50
    rr = R.KNRM(provide={"extractor": Extractor(synthetic.embedding_table(V, E, seed=0), Q, D)})
    model = rr.build_model().to(DEV).eval()
    b = {k: torch.from_numpy(v).to(DEV) for k, v in synthetic.parity_batch(6, Q, D, V, seed=3).items()}
    with torch.no_grad():
        s0 = rr.test(b).clone()
        model.embedding.weight.mul_(-1.0)  # cosines are invariant to a global sign flip
        s1 = rr.test(b).clone()
        model.embedding.weight[1:50].normal_()
        s2 = rr.test(b)
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-5)
    assert not torch.allclose(s0, s2, rtol=1e-3, atol=1e-3)

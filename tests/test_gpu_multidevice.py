"""GPU (>= 2 devices): every C entry point works on the device that OWNS its buffers, whatever the current device is (device guard
in the C ABI, capreolus_b200/csrc/common.cuh::DeviceGuard).  A model moved to cuda:1 while cuda:0 is current gives the bits it gives on
cuda:0 -- including PACRR, whose filter bank lives in per-device `__constant__` memory, and the BERT engine with its cached workspace.
Skipped on single-GPU boxes (run with `gpurun --gpus 2`)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_parity import DRMM_CFG, KNRM_CFG, PACRR_CFG, _batch, _build

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


@pytest.mark.parametrize("cls,cfgs", [("KNRM", KNRM_CFG), ("DRMM", DRMM_CFG), ("PACRR", PACRR_CFG)])
def test_scores_on_the_second_device_equal_the_first(cls, cfgs):
    g = load_golden(f"{cls.lower()}_small")
    rr, model = _build(cls, g, "default", cfgs["default"])
    with torch.no_grad():
        torch.cuda.set_device(0)
        first = rr.test(_batch(g, "cuda:0")).cpu().numpy()
        model.to("cuda:1")
        assert torch.cuda.current_device() == 0
        second = rr.test(_batch(g, "cuda:1"))  # current device is still cuda:0
        assert second.device == torch.device("cuda:1")
        torch.cuda.synchronize(1)
        again_first_dev = None
        model.to("cuda:0")
        again_first_dev = rr.test(_batch(g, "cuda:0")).cpu().numpy()
    assert np.array_equal(first, second.cpu().numpy())
    assert np.array_equal(first, again_first_dev)


def test_two_pacrr_models_on_two_devices_interleaved():
    """Different filter banks on the two devices, calls interleaved from one thread: each device's `__constant__` bank holds its own model."""
    g = load_golden("pacrr_small")
    rr_a, model_a = _build("PACRR", g, "default", PACRR_CFG["default"])
    rr_b, model_b = _build("PACRR", g, "default", PACRR_CFG["default"])
    with torch.no_grad():
        for p in model_b.parameters():
            p.mul_(1.25)
        want_a = rr_a.test(_batch(g, "cuda:0")).cpu().numpy()
        want_b = rr_b.test(_batch(g, "cuda:0")).cpu().numpy()
        assert not np.array_equal(want_a, want_b)
        model_b.to("cuda:1")
        ba, bb = _batch(g, "cuda:0"), _batch(g, "cuda:1")
        for _ in range(3):
            got_a = rr_a.test(ba)
            got_b = rr_b.test(bb)
        torch.cuda.synchronize(0), torch.cuda.synchronize(1)
    assert np.array_equal(got_a.cpu().numpy(), want_a)
    assert np.array_equal(got_b.cpu().numpy(), want_b)


def test_bert_engine_follows_the_model_to_another_device():
    from test_gpu_bert import _build as build_bert

    g, rr, model, b = build_bert("tiny")
    with torch.no_grad():
        torch.cuda.set_device(0)
        first = rr.test(b).cpu().numpy()
        model.to("cuda:1")
        b1 = {k: v.to("cuda:1") for k, v in b.items()}
        second = rr.test(b1)
        assert second.device == torch.device("cuda:1")
    assert np.array_equal(first, second.cpu().numpy())

"""ctypes binding of ``libcapr_b200.so`` (C ABI: ``include/capr_b200.h``).

This is the only place the Python host code touches native code.  There is no fallback: if the shared
library is missing (``python -c "import __graft_entry__ as g; g.build()"`` builds it) or a call fails,
an exception is raised -- scoring never silently runs on another path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_void_p
from pathlib import Path

LIB_NAME = "libcapr_b200.so"
LIB_PATH = Path(os.environ.get("CAPR_B200_LIB", Path(__file__).resolve().parent / LIB_NAME))
#: debug build of the same sources (-DCAPR_DEBUG_BUILD) + the micro-benchmarks; loaded only by dbg_lib() (tests, scripts, the
#: L2-gather roofline probe of bench.py).  Setting CAPR_B200_LIB to it makes the profiling switches (CAPR_DEBUG_FLAGS ...) live.
DBG_LIB_PATH = Path(os.environ.get("CAPR_B200_DBG_LIB", Path(__file__).resolve().parent / "libcapr_b200_dbg.so"))

CAPR_OK, BAD_SHAPE, BAD_POINTER, UNSUPPORTED, CUDA_ERROR, NO_DEVICE = 0, -1, -2, -3, -4, -5

_lib = None
_dbg_lib = None

_f32p, _i64p = c_void_p, c_void_p  # device pointers travel as plain addresses (tensor.data_ptr())


class BertConfigStruct(Structure):
    """``capr_bert_config`` of include/capr_b200.h"""

    _fields_ = [("hidden", c_int), ("layers", c_int), ("heads", c_int), ("intermediate", c_int), ("vocab", c_int), ("max_pos", c_int),
                ("type_vocab", c_int), ("n_labels", c_int), ("ln_eps", c_float)]


BERT_BF16, BERT_BF16X3 = 1, 3

#: every symbol include/capr_b200.h declares -> (restype, argtypes); tests check the .so exports them all
SIGNATURES = {
    "capr_abi_version": (c_int, []),
    "capr_last_error": (c_char_p, []),
    "capr_device_sm_count": (c_int, []),
    "capr_device_arch": (c_int, []),
    "capr_table_pitch": (c_int, [c_int]),
    "capr_table_prepare": (c_int, [_f32p, c_int, c_int, _f32p, c_int, c_void_p]),
    "capr_simmat_forward": (c_int, [_i64p, _i64p, c_int, c_int, c_int, _f32p, c_int, c_int, _f32p, c_void_p]),
    "capr_knrm_forward": (c_int, [_i64p, _i64p, c_int, c_int, c_int, _f32p, c_int, c_int, _f32p, _f32p, c_int, _f32p, _f32p,
                                  c_int, _f32p, _f32p, c_int, _f32p, _f32p, _f32p, c_void_p]),
    "capr_table_pitch_bf16": (c_int, [c_int]),
    "capr_table_prepare_bf16": (c_int, [_f32p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "capr_knrm_forward_tc": (c_int, [_i64p, _i64p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, _f32p, _f32p, c_int,
                                     _f32p, _f32p, c_int, _f32p, _f32p, c_int, _f32p, _f32p, c_void_p]),
    "capr_tf_workspace_bytes": (c_size_t, [c_int, c_int]),
    "capr_tf_dedup": (c_int, [_i64p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "capr_knrm_forward_tf": (c_int, [_i64p, _i64p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, _f32p, _f32p, c_int,
                                     _f32p, _f32p, c_int, _f32p, _f32p, c_int, _f32p, _f32p, c_void_p, c_size_t, c_void_p]),
    "capr_drmm_forward": (c_int, [_i64p, _i64p, _f32p, c_int, c_int, c_int, _f32p, c_int, c_int, _f32p, c_int, c_int, _f32p,
                                  c_int, c_int, _f32p, _f32p, c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, c_void_p]),
    "capr_drmm_forward_tc": (c_int, [_i64p, _i64p, _f32p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, _f32p, c_int, c_int, _f32p,
                                     c_int, c_int, _f32p, _f32p, c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, c_void_p]),
    "capr_pacrr_forward_tc": (c_int, [_i64p, _i64p, _f32p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, POINTER(c_void_p), POINTER(c_void_p), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, c_int, c_int,
                                      _f32p, _f32p, c_void_p]),
    "capr_pacrr_forward": (c_int, [_i64p, _i64p, _f32p, c_int, c_int, c_int, _f32p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   POINTER(c_void_p), POINTER(c_void_p), _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, c_int, c_int,
                                   _f32p, _f32p, c_void_p]),
    "capr_drmmtks_forward_tc": (c_int, [_i64p, _i64p, _f32p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, _f32p, _f32p,
                                        _f32p, _f32p, _f32p, _f32p, _f32p, c_void_p]),
    "capr_convknrm_proj_cols": (c_int, [c_int, c_int]),
    "capr_convknrm_project": (c_int, [_f32p, c_int, c_int, POINTER(c_void_p), c_int, c_int, _f32p, c_void_p]),
    "capr_convknrm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "capr_convknrm_forward": (c_int, [_i64p, _i64p, c_int, c_int, c_int, _f32p, c_int, c_int, c_int, POINTER(c_void_p), c_int, _f32p, _f32p, c_int,
                                      _f32p, _f32p, c_int, _f32p, _f32p, c_int, _f32p, _f32p, c_void_p, c_size_t, c_void_p]),
    "capr_assemble_pairs": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, _f32p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    _i64p, _i64p, _f32p, c_void_p]),
    "capr_assemble_bert_pairs": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, _i64p, _i64p, _i64p, c_void_p]),
    "capr_widen_ids": (c_int, [c_void_p, c_int, c_size_t, _i64p, c_void_p]),
    "capr_rank_by_query": (c_int, [_f32p, c_void_p, c_int, c_int, _f32p, c_void_p, c_void_p]),
    "capr_pair_hinge": (c_int, [_f32p, _f32p, c_int, _f32p, _f32p, _f32p, c_void_p]),
    "capr_pair_softmax": (c_int, [_f32p, _f32p, c_int, _f32p, _f32p, _f32p, c_void_p]),
    "capr_bert_num_weights": (c_int, [POINTER(BertConfigStruct)]),
    "capr_bert_create": (c_int, [POINTER(BertConfigStruct), POINTER(c_void_p), c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "capr_bert_destroy": (None, [c_void_p]),
    "capr_bert_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "capr_bert_forward": (c_int, [c_void_p, _i64p, _i64p, _i64p, c_int, c_int, _f32p, c_void_p, c_size_t, c_void_p]),
    "capr_bert_forward_hidden": (c_int, [c_void_p, _i64p, _i64p, _i64p, c_int, c_int, POINTER(c_int), c_int, _f32p, _f32p, c_void_p, c_size_t, c_void_p]),
    "capr_cedrknrm_feature_dim": (c_int, [c_int, c_int, c_int, c_int]),
    "capr_cedrknrm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "capr_cedrknrm_head": (c_int, [_f32p, c_int, _f32p, _i64p, _i64p, c_int, c_int, c_int, c_int, c_int, _f32p, _f32p, c_int, c_int, _f32p, _f32p,
                                   c_int, _f32p, _f32p, _f32p, _f32p, c_void_p, c_size_t, c_void_p]),
    "capr_parade_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "capr_parade_head": (c_int, [c_void_p, _f32p, c_int, c_int, c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, c_void_p, c_size_t, c_void_p]),
}

#: exported by libcapr_b200_dbg.so only (include/capr_b200.h, `#ifdef CAPR_DEBUG_BUILD`)
DEBUG_SIGNATURES = {
    "capr_debug_mma_bench": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "capr_debug_ffma2_bench": (c_int, [c_int, c_int, c_int, _f32p, c_void_p, c_void_p]),
    "capr_gemm_test": (c_int, [_f32p, _f32p, _f32p, c_int, c_int, c_int, c_int, _f32p, c_void_p]),
    "capr_debug_gather_bench": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "capr_debug_gather_bench2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "capr_debug_gather_bench3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
}


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: the CUDA kernels are not built. Run `make` (or `python -c 'import __graft_entry__ as g; "
                f"g.build()'`) in the repo root. capreolus_b200 has no CPU or PyTorch fallback."
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = restype, argtypes
        if handle.capr_abi_version() != 1:
            raise RuntimeError(f"{LIB_PATH}: ABI version {handle.capr_abi_version()} != 1 (stale build?)")
        _lib = handle
    return _lib


def dbg_lib() -> ctypes.CDLL:
    """The debug build (profiling switches, ``capr_gemm_test``, ``capr_debug_*`` micro-benchmarks).  Never used for scoring."""
    global _dbg_lib
    if _dbg_lib is None:
        if not DBG_LIB_PATH.exists():
            raise NativeLibraryMissing(f"{DBG_LIB_PATH} not found: run `make` in the repo root")
        handle = ctypes.CDLL(str(DBG_LIB_PATH))
        for name, (restype, argtypes) in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _dbg_lib = handle
    return _dbg_lib


def check(rc: int, handle=None) -> None:
    """Map a capr_status to the exception type the reference would raise (SURVEY.md §8b 'Errors')."""
    if rc == CAPR_OK:
        return
    msg = ((handle or lib()).capr_last_error() or b"").decode("utf-8", "replace")
    if rc in (BAD_SHAPE, UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(f"capr_b200 error {rc}: {msg}")


def ptr(t) -> int | None:
    """Raw device address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "capreolus_b200 scores on a B200 through its CUDA kernels only; got a CPU tensor. Move the model and the batch "
                "to cuda (the reference trainer does: trainer/pytorch.py:94,203,328,342). There is no CPU fallback."
            )

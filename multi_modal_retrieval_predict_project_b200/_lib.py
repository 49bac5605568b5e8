"""ctypes binding of libmmr_b200.so (the C ABI in include/mmr_b200.h).

PyTorch supplies device memory and streams; this module only passes raw pointers.  There is no
CPU fallback: if the library is missing or no sm_100 device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMR_B200_LIB") or os.path.join(_HERE, "libmmr_b200.so")  # override: A/B builds

MMR_OK, MMR_EINVAL, MMR_ECUDA, MMR_ENOMEM, MMR_ENODEV, MMR_EUNSUP = range(6)
MMR_F32, MMR_BF16 = 0, 1
ALGO_AUTO, ALGO_SCAN, ALGO_GEMM = 0, 1, 2
FLAG_BORROW = 1
MAX_K = 1024
ABI_VERSION = 2

TUNE_GEMM_VARIANT, TUNE_GEMM_PARTS, TUNE_GEMM_PAIR = 0, 1, 2
GEMM_VARIANTS = {"auto": 0, "long": 1, "short": 2}

ALGOS = {"auto": ALGO_AUTO, "scan": ALGO_SCAN, "gemm": ALGO_GEMM}

_lib = None

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "mmr_abi_version": [],
    "mmr_last_error": [],
    "mmr_launch_count": [],
    "mmr_index_profile": [_vp, _i32, C.POINTER(_f64), C.POINTER(_i32)],
    "mmr_index_tune": [_vp, _i32, _i32],
    "mmr_index_last_plan": [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)],
    "mmr_index_create": [C.POINTER(_vp), _vp, _i64, _i32, _i32, _i32, _i64, _i32, _i32, _vp],
    "mmr_index_destroy": [_vp],
    "mmr_index_info": [_vp, C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32),
                       C.POINTER(_i64), C.POINTER(_i64)],
    "mmr_index_device_ptrs": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
    "mmr_index_get_rows": [_vp, _vp, _i64, _vp, _vp],
    "mmr_search": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "mmr_merge_topk": [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp],
    "mmr_merge_topk_strided": [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _vp],
    "mmr_gather_payload": [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _i32, _vp],
    "mmr_apply_order": [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "mmr_exchange_create": [C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32],
    "mmr_exchange_handle_bytes": [],
    "mmr_exchange_handle": [_vp, _vp],
    "mmr_exchange_open": [_vp, _vp],
    "mmr_exchange_destroy": [_vp],
    "mmr_exchange_set_timeout": [_vp, _i32],
    "mmr_exchange_status": [_vp, C.POINTER(_i32)],
    "mmr_exchange_abort": [_vp],
    "mmr_exchange_close_peers": [_vp],
    "mmr_search_scatter": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, C.c_uint32, _vp],
    "mmr_exchange_rerank": [_vp, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _i32, C.c_uint32, C.POINTER(_vp), C.POINTER(_vp), _vp],
    "mmr_rerank_scored": [_vp, _vp, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _i32, _vp, _vp, _vp, _i32, _vp],
    "mmr_candidate_cosine": [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp],
    "mmr_rerank_with_cos": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _i32, _vp, _vp, _i32, _vp],
    "mmr_rerank_tables_create": [C.POINTER(_vp), _vp, _i32, _vp, _i32, _i64, _i32, _vp],
    "mmr_rerank_tables_destroy": [_vp],
    "mmr_rerank_features": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp],
    "mmr_rerank_combine": [_vp, _vp, _i32, _i32, _f64, _f64, _f64, _i32, _vp, _vp, _i32, _vp],
    "mmr_rerank": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f64, _f64, _f64, _i32, _vp, _vp, _vp],
    "mmr_metrics": [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp],
    "mmr_first_relevant_rank": [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp],
    "mmr_label_ranking_eval": [_vp, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp],
    "mmr_result_diversity": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "mmr_label_relevance": [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, _vp],
}
_RESTYPES = {"mmr_last_error": C.c_char_p, "mmr_launch_count": C.c_int64}


class MMRError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libmmr_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m multi_modal_retrieval_predict_project_b200.build` "
            "(needs nvcc).  The B200 retrieval path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI symbol missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    got = lib.mmr_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libmmr_b200.so ABI version {got} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == MMR_OK:
        return
    msg = load().mmr_last_error()
    msg = msg.decode("utf8", "replace") if msg else f"libmmr_b200 error {status}"
    if status == MMR_EINVAL:
        raise ValueError(msg)
    if status == MMR_ENOMEM:
        raise MemoryError(msg)
    if status == MMR_EUNSUP:
        raise NotImplementedError(msg)
    raise MMRError(msg)


def ptr(x) -> Optional[int]:
    """Raw address of a numpy array / torch tensor (must be contiguous), or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(f"cannot take the address of {type(x)!r}")


class _DevMem:
    """__cuda_array_interface__ view of library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def as_cuda_tensor(ptr: int, shape, dtype, device: int):
    """Zero-copy torch view of device memory owned by the library (valid as long as the owner says)."""
    import torch
    typestr = {torch.int64: "<i8", torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_DevMem(ptr, shape, typestr), device=torch.device("cuda", device))


def current_stream(device: int) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device: Optional[int] = None) -> int:
    """Pick the CUDA device ordinal; raise loudly when there is none (no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available: the B200 retrieval path has no CPU fallback")
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, str):
        return torch.device(device).index or 0
    if hasattr(device, "index"):
        return device.index or 0
    return int(device)

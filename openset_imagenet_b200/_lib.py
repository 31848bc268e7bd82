"""ctypes binding of libosb.so — the only way Python reaches the CUDA kernels.

The signatures below mirror ``include/osb.h`` one to one (a CPU test checks that every symbol the
header declares is exported and bound).  There is no fallback: a missing library, a missing GPU or
a non-sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import _build

ABI_VERSION = 4
OK, ERR_ARG, ERR_SHAPE, ERR_WORKSPACE, ERR_CUDA, ERR_ARCH = 0, -1, -2, -3, -4, -5
ENTROPIC, SOFTMAX, GARBAGE = 0, 1, 2
F32, BF16 = 0, 1
LABEL_I64, LABEL_F32, LABEL_I32 = 0, 1, 2
FWD_DEN_FROM_PREV_KERNEL = 1
FWD_PUBLISH_ROW_TILES = 2
BWD_DZ_FROM_PREV_FORWARD = 1

_p, _i, _i64, _sz, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); kept in header order
SIGNATURES = {
    "osb_abi_version": (_i, []),
    "osb_last_error": (C.c_char_p, []),
    "osb_device_check": (_i, []),
    "osb_sm_count": (_i, []),
    "osb_label_stats": (_i, [_p, _p, _i64, _i, _i, _i64, _p, _p, _p, _sz, _i, _p]),
    "osb_loss_workspace_bytes": (_sz, [_i64]),
    "osb_loss_from_logits": (_i, [_p, _i, _i64, _p, _p, _i64, _i, _i, _d, _i64, _p, _i,
                                  _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _sz, _p]),
    "osb_head_workspace_bytes": (_sz, [_i64, _i, _i]),
    "osb_head_forward": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _i, _i, _i, _d, _i64, _p, _i,
                              _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _sz, _p]),
    "osb_head_backward": (_i, [_p, _i64, _i, _p, _p, _i64, _p, _i64, _i, _i, _p, _p, _i64, _i, _p, _p, _i, _i, _p, _sz, _p]),
    "osb_head_forward_dx": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _i, _i, _i, _d, _i64, _p, _i,
                                 _p, _p, _p, _p, _p, _i64, _i, _p, _p, _i64, _i, _i, _i, _p, _sz, _p]),
    "osb_head_forward_dx_is_fused": (_i, [_i64, _i, _i]),
    "osb_head_forward_is_splitk": (_i, [_i64, _i, _i]),
    "osb_head_dw_is_direct": (_i, [_i64, _i, _i]),
    "osb_head_dw_partials_offset": (_sz, []),
    "osb_head_dw_splits": (_i, [_i64, _i, _i]),
    "osb_head_backward_dw_partials": (_i, [_p, _i64, _i, _p, _i64, _i64, _i, _i, _p, _i, _i, _p, _sz, _p]),
    "osb_head_dw_allreduce_flag_bytes": (_sz, [_i]),
    "osb_head_dw_allreduce": (_i, [_p, _p, _p, _p, _i, _i, _i, _i64, _i64, _p, _i, _i, _p, _p, _p, _p, _i, _p]),
    "osb_head_dw_allreduce_error": (_i, [_p, _i, _p]),
    "osb_peer_sum_bytes": (_sz, [_i]),
    "osb_peer_sum": (_i, [_p, _p, _i, _p, _i, _i, _p]),
    "osb_confidence_workspace_bytes": (_sz, [_i64]),
    "osb_confidence": (_i, [_p, _i64, _p, _i64, _i, _f, _i64, _i, _p, _p, _p, _sz, _p]),
    "osb_confidence_from_stats": (_i, [_p, _p, _p, _i64, _i, _f, _i64, _i, _p, _p, _p]),
    "osb_oscr_workspace_bytes": (_sz, [_i64]),
    "osb_oscr_curve_workspace_bytes": (_sz, [_i64]),
    "osb_oscr_records": (_i, [_p, _i64, _p, _i, _i64, _i, _i64, _p, _p, _p, _p]),
    "osb_oscr_records_from_stats": (_i, [_p, _p, _p, _p, _i, _i64, _i, _i64, _p, _p, _p, _p]),
    "osb_oscr_curve": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "osb_oscr_ratios": (_i, [_p, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "osb_oscr": (_i, [_p, _i64, _p, _i, _i64, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "osb_oscr_with_confidence": (_i, [_p, _i64, _p, _i, _i64, _i, _i64, _f, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "osb_oscr_exchange_slot_bytes": (_sz, [_i]),
    "osb_oscr_exchange": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i64, _p, _i, _p]),
    "osb_ccr_at_fpr": (_i, [_p, _p, _p, _i64, _p, _i, _p, _p, _p, _p]),
    "osb_hist_metrics": (_i, [_p, _i64, _p, _i64, _i, _p, _i64, _i, _i, _i64, _p, _p, _p, _p, _p]),
    "osb_histogram": (_i, [_p, _i64, _p, _i, _p, _p]),
}

_lock = threading.Lock()
_lib = None


class OsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libosb error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """dlopen libosb.so (building it first if the .so is absent and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            _build.build()          # raises loudly when nvcc is missing: there is no other path
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if lib.osb_abi_version() != ABI_VERSION:
            raise RuntimeError("libosb ABI version mismatch; rebuild with python -m openset_imagenet_b200._build --force")
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != OK:
        raise OsbError(rc, lib.osb_last_error().decode(errors="replace"))


def size(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """The product path is CUDA-only; a CPU tensor is a caller error, never a silent fallback."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("openset_imagenet_b200 runs on CUDA (sm_100a) tensors only; got a "
                               f"{t.device} tensor and there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_ws_cache: dict[tuple, torch.Tensor] = {}


def workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    """Zero-initialised, cached per (kind, device, stream); grown geometrically."""
    key = (kind, device.index, stream_ptr(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes * 1.25), 4096), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf

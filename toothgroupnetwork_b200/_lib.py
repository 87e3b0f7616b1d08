"""ctypes binding of libtgn_b200.so (include/tgn_b200.h).

The library is the product: if it is missing or a call fails this module raises -- there is no
CPU or PyTorch fallback.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libtgn_b200.so")
_lib: Optional[ctypes.CDLL] = None

_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float

# name -> argument types of the status-returning (part 2) entry points
_SIGS = {
    "tgn_furthestsampling": [_i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "tgn_knnquery": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "tgn_grouping_forward": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "tgn_grouping_backward": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "tgn_interpolation_forward": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_interpolation_backward": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_subtraction_forward": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_subtraction_backward": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_aggregation_forward": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "tgn_aggregation_backward": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "tgn_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _i, _vp],
    "tgn_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "tgn_gather_rows": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "tgn_transpose_cn": [_i, _i, _i, _vp, _vp, _vp],
    "tgn_sa_group_mlp_max": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
}

# the reference's ten extern "C" launchers (part 1): void return, legacy default stream
REFERENCE_LAUNCHERS = [
    "furthestsampling_cuda_launcher", "knnquery_cuda_launcher",
    "grouping_forward_cuda_launcher", "grouping_backward_cuda_launcher",
    "interpolation_forward_cuda_launcher", "interpolation_backward_cuda_launcher",
    "subtraction_forward_cuda_launcher", "subtraction_backward_cuda_launcher",
    "aggregation_forward_cuda_launcher", "aggregation_backward_cuda_launcher",
]
EXPORTS = list(_SIGS) + REFERENCE_LAUNCHERS + ["tgn_version", "tgn_last_error", "tgn_launch_count"]


class TgnError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TgnError(
            f"{LIB_PATH} is missing: build it with `python -m toothgroupnetwork_b200.build` "
            "(nvcc, sm_100a).  toothgroupnetwork_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    lib.tgn_last_error.restype = ctypes.c_char_p
    lib.tgn_version.restype = _i
    lib.tgn_launch_count.restype = _i
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load().tgn_launch_count())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def call(name: str, *args) -> None:
    lib = load()
    status = getattr(lib, name)(*args)
    if status != 0:
        raise TgnError(f"{name} failed ({status}): {lib.tgn_last_error().decode()}")


def require_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TgnError("toothgroupnetwork_b200 operators need CUDA tensors (no CPU path)")

"""Stand-in for the reference's pybind module ``pointops_cuda``
(``external_libs/pointops/src/pointops_api.cpp:12-23``): the same ten functions, same argument
order and the same caller-owns-every-buffer convention (outputs that are accumulated into must
be pre-zeroed, ``tmp`` pre-filled with 1e10), so the reference's own ``pointops.py`` runs on top
of it unchanged.  Each call forwards raw device pointers to libtgn_b200.so on the current
PyTorch stream (the reference used the legacy default stream).
"""
from __future__ import annotations

import torch

from . import _lib as L


def _chk_f32(*ts):
    for t in ts:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise L.TgnError("expected a contiguous float32 tensor")


def _chk_i32(*ts):
    for t in ts:
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise L.TgnError("expected a contiguous int32 tensor")


def furthestsampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    """sampling/sampling_cuda.cpp:8-16.  ``n`` is the largest cloud size."""
    L.require_cuda(xyz, offset, new_offset, tmp, idx)
    _chk_f32(xyz, tmp)
    _chk_i32(offset, new_offset, idx)
    L.call("tgn_furthestsampling", int(b), int(n), L.ptr(xyz), L.ptr(offset), L.ptr(new_offset), L.ptr(tmp), L.ptr(idx),
           0, L.stream_ptr())


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """knnquery/knnquery_cuda.cpp:8-17."""
    L.require_cuda(xyz, new_xyz, offset, new_offset, idx, dist2)
    _chk_f32(xyz, new_xyz, dist2)
    _chk_i32(offset, new_offset, idx)
    L.call("tgn_knnquery", int(offset.shape[0]), int(m), int(nsample), L.ptr(xyz), L.ptr(new_xyz), L.ptr(offset),
           L.ptr(new_offset), L.ptr(idx), L.ptr(dist2), L.stream_ptr())


def grouping_forward_cuda(m, nsample, c, input, idx, output):
    L.require_cuda(input, idx, output)
    _chk_f32(input, output)
    _chk_i32(idx)
    L.call("tgn_grouping_forward", int(m), int(nsample), int(c), L.ptr(input), L.ptr(idx), L.ptr(output), L.stream_ptr())


def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
    L.require_cuda(grad_output, idx, grad_input)
    _chk_f32(grad_output, grad_input)
    _chk_i32(idx)
    L.call("tgn_grouping_backward", int(m), int(nsample), int(c), L.ptr(grad_output), L.ptr(idx), L.ptr(grad_input), L.stream_ptr())


def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
    L.require_cuda(input, idx, weight, output)
    _chk_f32(input, weight, output)
    _chk_i32(idx)
    L.call("tgn_interpolation_forward", int(n), int(c), int(k), L.ptr(input), L.ptr(idx), L.ptr(weight), L.ptr(output), L.stream_ptr())


def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
    L.require_cuda(grad_output, idx, weight, grad_input)
    _chk_f32(grad_output, weight, grad_input)
    _chk_i32(idx)
    L.call("tgn_interpolation_backward", int(n), int(c), int(k), L.ptr(grad_output), L.ptr(idx), L.ptr(weight), L.ptr(grad_input), L.stream_ptr())


def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
    L.require_cuda(input1, input2, idx, output)
    _chk_f32(input1, input2, output)
    _chk_i32(idx)
    L.call("tgn_subtraction_forward", int(n), int(nsample), int(c), L.ptr(input1), L.ptr(input2), L.ptr(idx), L.ptr(output), L.stream_ptr())


def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
    L.require_cuda(idx, grad_output, grad_input1, grad_input2)
    _chk_f32(grad_output, grad_input1, grad_input2)
    _chk_i32(idx)
    L.call("tgn_subtraction_backward", int(n), int(nsample), int(c), L.ptr(idx), L.ptr(grad_output), L.ptr(grad_input1), L.ptr(grad_input2), L.stream_ptr())


def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
    L.require_cuda(input, position, weight, idx, output)
    _chk_f32(input, position, weight, output)
    _chk_i32(idx)
    L.call("tgn_aggregation_forward", int(n), int(nsample), int(c), int(w_c), L.ptr(input), L.ptr(position), L.ptr(weight),
           L.ptr(idx), L.ptr(output), L.stream_ptr())


def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input, grad_position, grad_weight):
    L.require_cuda(input, position, weight, idx, grad_output, grad_input, grad_position, grad_weight)
    _chk_f32(input, position, weight, grad_output, grad_input, grad_position, grad_weight)
    _chk_i32(idx)
    L.call("tgn_aggregation_backward", int(n), int(nsample), int(c), int(w_c), L.ptr(input), L.ptr(position), L.ptr(weight),
           L.ptr(idx), L.ptr(grad_output), L.ptr(grad_input), L.ptr(grad_position), L.ptr(grad_weight), L.stream_ptr())

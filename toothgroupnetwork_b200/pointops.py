"""Operator API of ``external_libs/pointops/functions/pointops.py`` on libtgn_b200.so.

Same public names, positional signatures, dtypes and layouts as the reference (packed
``(n_total, C)`` tensors with cumulative int32 ``offset``), so the callers in
``models/modules/cbl_point_transformer/{blocks,heads,basic_operators}.py`` and ``gen_utils.fps``
run unchanged.  Every op launches hand-written sm_100a kernels on the current stream; there is
no CPU path.  Reference line numbers below refer to that file.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib as L

_RESIDENT_FPS_MAX_POINTS = 8 * 12288   # largest cloud the register-resident FPS clusters hold


def _empty(shape, dtype, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=like.device)


def _need(t: torch.Tensor, dtype, what: str) -> None:
    if not t.is_cuda:
        raise L.TgnError(f"{what}: CUDA tensor required (no CPU path)")
    if t.dtype != dtype:
        raise L.TgnError(f"{what}: expected {dtype}, got {t.dtype}")
    assert t.is_contiguous(), f"{what} must be contiguous"   # the reference asserts the same (:17,:38,...)


def fps_packed(xyz: torch.Tensor, offset: torch.Tensor, new_offset: torch.Tensor, n_max: int, m_total: int,
               mode: int = 0) -> torch.Tensor:
    """FPS with the host-side sizes already known (no device->host sync).  ``mode``: 0 auto,
    -1 the streaming kernel, otherwise 100*G + CS forces a cluster of CS CTAs with G clouds in
    flight (G omitted = 1), e.g. 8 or 204."""
    _need(xyz, torch.float32, "xyz")
    _need(offset, torch.int32, "offset")
    _need(new_offset, torch.int32, "new_offset")
    idx = torch.zeros((m_total,), dtype=torch.int32, device=xyz.device)   # zero-filled like the reference (:21)
    if m_total == 0:
        return idx
    tmp = None
    if mode == -1 or n_max > _RESIDENT_FPS_MAX_POINTS:
        tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)   # :22
    L.call("tgn_furthestsampling", int(offset.shape[0]), int(n_max), L.ptr(xyz), L.ptr(offset), L.ptr(new_offset),
           L.ptr(tmp), L.ptr(idx), int(mode), L.stream_ptr())
    return idx


def _offset_sizes(offset: torch.Tensor, new_offset: torch.Tensor, n_total: int) -> Tuple[int, int]:
    """(largest segment, total samples) of an (offset, new_offset) pair: ONE host copy of both vectors, where the
    reference reads them back with one blocking ``.item()`` per cloud (:18-21).  The output size is data, so one
    read-back is the floor for this signature; its callers (``blocks.py:64-69``) have just synchronised on
    ``o[i].item()`` themselves.  Callers that know the sizes use ``fps_packed`` directly and never synchronise."""
    host = torch.stack([offset.to(torch.int64), new_offset.to(torch.int64)]).cpu()
    n_max = int(torch.diff(host[0], prepend=host[0].new_zeros(1)).max())
    return n_max, int(host[1][-1])


class FurthestSampling(Function):
    """:10-27.  input: xyz (n,3), offset (b), new_offset (b); output: idx (m) int32 global row ids."""

    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        if offset.shape[0] == 0:
            return _empty((0,), torch.int32, xyz)
        n_max, m_total = _offset_sizes(offset, new_offset, xyz.shape[0])
        idx = fps_packed(xyz, offset, new_offset, n_max, m_total)
        ctx.mark_non_differentiable(idx)
        return idx


furthestsampling = FurthestSampling.apply


def knn_packed(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, offset: torch.Tensor,
               new_offset: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx (m,nsample) int32 and SQUARED distances (m,nsample)."""
    _need(xyz, torch.float32, "xyz")
    _need(new_xyz, torch.float32, "new_xyz")
    _need(offset, torch.int32, "offset")
    _need(new_offset, torch.int32, "new_offset")
    m = new_xyz.shape[0]
    idx = _empty((m, nsample), torch.int32, xyz)
    d2 = _empty((m, nsample), torch.float32, xyz)
    L.call("tgn_knnquery", int(offset.shape[0]), int(m), int(nsample), L.ptr(xyz), L.ptr(new_xyz), L.ptr(offset),
           L.ptr(new_offset), L.ptr(idx), L.ptr(d2), L.stream_ptr())
    return idx, d2


class KNNQuery(Function):
    """:30-45.  output: idx (m,nsample) int32, dist (m,nsample) = sqrt of the squared distances."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz, offset, new_offset):
        if new_xyz is None:
            new_xyz = xyz
        idx, d2 = knn_packed(int(nsample), xyz, new_xyz, offset, new_offset)
        dist = torch.sqrt(d2)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist


knnquery = KNNQuery.apply


class Grouping(Function):
    """:48-76.  input (n,c), idx (m,nsample) -> (m,nsample,c); backward scatter-adds."""

    @staticmethod
    def forward(ctx, input, idx):
        _need(input, torch.float32, "input")
        _need(idx, torch.int32, "idx")
        m, nsample = idx.shape
        n, c = input.shape
        out = _empty((m, nsample, c), torch.float32, input)
        L.call("tgn_grouping_forward", m, nsample, c, L.ptr(input), L.ptr(idx), L.ptr(out), L.stream_ptr())
        ctx.n = n
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        m, nsample, c = grad_output.shape
        grad_in = torch.zeros((ctx.n, c), dtype=torch.float32, device=grad_output.device)
        L.call("tgn_grouping_backward", m, nsample, c, L.ptr(grad_output), L.ptr(idx), L.ptr(grad_in), L.stream_ptr())
        return grad_in, None


grouping = Grouping.apply


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """:79-100.  kNN (unless idx is given) + gather + centre-subtract + concat.
    output (m, nsample, 3+c) with channel order [xyz_rel, feat], or (m, nsample, c)."""
    assert xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    assert new_xyz.is_contiguous()
    if idx is None:
        idx, _ = knnquery(nsample, xyz, new_xyz, offset, new_offset)
    grouped_feat = grouping(feat, idx)
    if not use_xyz:
        return grouped_feat
    grouped_xyz = grouping(xyz, idx) - new_xyz.unsqueeze(1)
    return torch.cat((grouped_xyz, grouped_feat), -1)


class Subtraction(Function):
    """:103-130.  out[n,s,:] = input1[n,:] - input2[idx[n,s],:]."""

    @staticmethod
    def forward(ctx, input1, input2, idx):
        _need(input1, torch.float32, "input1")
        _need(input2, torch.float32, "input2")
        _need(idx, torch.int32, "idx")
        n, c = input1.shape
        nsample = idx.shape[-1]
        out = _empty((n, nsample, c), torch.float32, input1)
        L.call("tgn_subtraction_forward", n, nsample, c, L.ptr(input1), L.ptr(input2), L.ptr(idx), L.ptr(out), L.stream_ptr())
        ctx.n2 = input2.shape[0]
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, nsample, c = grad_output.shape
        g1 = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
        g2 = torch.zeros((ctx.n2, c), dtype=torch.float32, device=grad_output.device)
        L.call("tgn_subtraction_backward", n, nsample, c, L.ptr(idx), L.ptr(grad_output), L.ptr(g1), L.ptr(g2), L.stream_ptr())
        return g1, g2, None


subtraction = Subtraction.apply


class Aggregation(Function):
    """:133-161.  out[n,c] = sum_s (input[idx[n,s],c] + position[n,s,c]) * weight[n,s,c % w_c]."""

    @staticmethod
    def forward(ctx, input, position, weight, idx):
        _need(input, torch.float32, "input")
        _need(position, torch.float32, "position")
        _need(weight, torch.float32, "weight")
        _need(idx, torch.int32, "idx")
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        out = torch.zeros((n, c), dtype=torch.float32, device=input.device)
        L.call("tgn_aggregation_forward", n, nsample, c, w_c, L.ptr(input), L.ptr(position), L.ptr(weight), L.ptr(idx),
               L.ptr(out), L.stream_ptr())
        ctx.save_for_backward(input, position, weight, idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, position, weight, idx = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        gi = torch.zeros_like(input)
        gp = torch.zeros_like(position)
        gw = torch.zeros_like(weight)
        L.call("tgn_aggregation_backward", n, nsample, c, w_c, L.ptr(input), L.ptr(position), L.ptr(weight), L.ptr(idx),
               L.ptr(grad_output), L.ptr(gi), L.ptr(gp), L.ptr(gw), L.stream_ptr())
        return gi, gp, gw, None


aggregation = Aggregation.apply


def _inverse_distance_weights(dist: torch.Tensor) -> torch.Tensor:
    """:171-173 / :192-194: 1/(dist+1e-8), normalised over the k neighbours."""
    rec = 1.0 / (dist + 1e-8)
    return rec / torch.sum(rec, dim=1, keepdim=True)


class _WeightedGather(Function):
    """out[n,:] = sum_i input[idx[n,i],:] * weight[n,i] with a scatter-add backward wrt input."""

    @staticmethod
    def forward(ctx, input, idx, weight, fused=True):
        _need(input, torch.float32, "input")
        _need(idx, torch.int32, "idx")
        _need(weight, torch.float32, "weight")
        n, k = idx.shape
        m, c = input.shape
        out = torch.zeros((n, c), dtype=torch.float32, device=input.device)
        L.call("tgn_weighted_gather", n, c, k, L.ptr(input), L.ptr(idx), L.ptr(weight), L.ptr(out), 1 if fused else 0, L.stream_ptr())
        ctx.m = m
        ctx.save_for_backward(idx, weight)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, c = grad_output.shape
        gi = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad_output.device)
        L.call("tgn_interpolation_backward", n, c, idx.shape[1], L.ptr(grad_output), L.ptr(idx), L.ptr(weight), L.ptr(gi), L.stream_ptr())
        return gi, None, None, None


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """:164-180.  input: xyz (m,3) coarse, new_xyz (n,3) fine, feat (m,c) -> (n,c).
    Weights are constants for autograd (the reference detaches them, :175)."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    weight = _inverse_distance_weights(dist).detach().contiguous()
    return _WeightedGather.apply(feat, idx, weight, False)      # the reference accumulates with torch ops here: unfused (:177-179)


class Interpolation(Function):
    """:183-216 (exported as ``interpolation2``; fused kernel form of the same operation)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, input, offset, new_offset, k=3):
        assert xyz.is_contiguous() and new_xyz.is_contiguous() and input.is_contiguous()
        idx, d2 = knn_packed(int(k), xyz, new_xyz, offset, new_offset)
        weight = _inverse_distance_weights(torch.sqrt(d2)).contiguous()
        n, c, m = new_xyz.shape[0], input.shape[1], input.shape[0]
        out = torch.zeros((n, c), dtype=torch.float32, device=input.device)
        L.call("tgn_interpolation_forward", n, c, int(k), L.ptr(input), L.ptr(idx), L.ptr(weight), L.ptr(out), L.stream_ptr())
        ctx.m, ctx.k = m, int(k)
        ctx.save_for_backward(idx, weight)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, c = grad_output.shape
        gi = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad_output.device)
        L.call("tgn_interpolation_backward", n, c, ctx.k, L.ptr(grad_output), L.ptr(idx), L.ptr(weight), L.ptr(gi), L.stream_ptr())
        return None, None, gi, None, None, None


interpolation2 = Interpolation.apply

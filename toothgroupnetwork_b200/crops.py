"""Crop extraction between the two stages of the reference's two-stage models, on the GPU.

The reference copies the cloud to the host, builds a sklearn ``KDTree`` and asks it for the ``crop_sample_size`` = 3072
nearest vertices of every tooth centroid (``ops_utils.get_nearest_neighbor_idx`` ops_utils.py:146-161, called at
``grouping_network_module.py:71-73`` and ``tsegnet.py:73``), gathers the crops back on the GPU
(``ops_utils.get_indexed_features`` :198-218) and centres them (``centering_object`` :164-169).  Here the selection is one
CUDA kernel (csrc/crop_knn.cu: radix select + bitonic sort per centre, float64 distances like the KDTree, ascending
order), the gather is an index on the device, and nothing leaves the GPU.

``accelerate(ops_utils_module)`` swaps the three functions of an imported reference ``ops_utils`` for these, keeping their
signatures (numpy or tensors in, the same containers out), so ``models/modules/*.py`` run unchanged.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np
import torch

from . import _lib as L

MAX_CROP = 4096


def nearest_neighbor_crops(xyz: torch.Tensor, centres: torch.Tensor, k: int) -> torch.Tensor:
    """xyz (B,N,3) float32 CUDA, centres (B,Q,3) -> (B,Q,k) int64: the k nearest points of every centre, ascending
    distance (ties: lower index first)."""
    L.require_cuda(xyz, centres)
    xyz = xyz.contiguous().float()
    centres = centres.to(xyz.device).contiguous().float()
    B, N, _ = xyz.shape
    Q = centres.shape[1]
    if k > MAX_CROP:
        raise L.TgnError(f"crop size {k} exceeds {MAX_CROP}")
    out = torch.empty((B, Q, k), dtype=torch.int64, device=xyz.device)
    L.call("tgn_crop_knn", B, N, Q, int(k), L.ptr(xyz), L.ptr(centres), L.ptr(out), 1, L.stream_ptr())
    return out


def get_nearest_neighbor_idx(org_xyz, sampled_clusters, crop_num: int = 4096) -> List:
    """Drop-in for ops_utils.get_nearest_neighbor_idx (:146-161): org_xyz (B,N,3) numpy or tensor, sampled_clusters a
    per-batch sequence of (Q_b,3) centres -> list of (Q_b, crop_num) index arrays (numpy in -> numpy out, tensor in ->
    CUDA tensors out)."""
    as_numpy = isinstance(org_xyz, np.ndarray)
    xyz = torch.as_tensor(org_xyz, dtype=torch.float32)
    xyz = xyz.cuda() if not xyz.is_cuda else xyz
    out = []
    for b in range(xyz.shape[0]):
        c = torch.as_tensor(np.asarray(sampled_clusters[b], dtype=np.float32) if not torch.is_tensor(sampled_clusters[b]) else sampled_clusters[b],
                            dtype=torch.float32).reshape(1, -1, 3).cuda()
        idx = nearest_neighbor_crops(xyz[b:b + 1], c, int(crop_num))[0]
        out.append(idx.cpu().numpy() if as_numpy else idx)
    return out


def get_indexed_features(features, cropped_indexes):
    """Drop-in for ops_utils.get_indexed_features (:198-218): features (B,C,N), cropped_indexes per batch (Q_b, K) ->
    (sum_b Q_b, C, K); index tensors already on the device are used as they are."""
    items = []
    for b in range(len(cropped_indexes)):
        idx = cropped_indexes[b]
        if torch.is_tensor(features):
            idx_t = idx if torch.is_tensor(idx) else torch.as_tensor(np.asarray(idx), device=features.device)
            items.append(features[b][:, idx_t.long()].permute(1, 0, 2).contiguous())     # (Q_b, C, K)
        else:
            items.append(np.stack([features[b][:, np.asarray(i)] for i in idx], axis=0))
    if torch.is_tensor(features):
        return torch.cat(items, dim=0)
    return np.concatenate(items, axis=0)


def accelerate(ops_utils_module) -> None:
    """Replace the KDTree crop search (and the gather) of an imported reference ``ops_utils`` module with the GPU versions."""
    ops_utils_module.get_nearest_neighbor_idx = get_nearest_neighbor_idx
    ops_utils_module.get_indexed_features = get_indexed_features


def nearest_label_transfer(src_xyz: torch.Tensor, src_labels: torch.Tensor, dst_xyz: torch.Tensor) -> torch.Tensor:
    """Labels of the nearest source vertex for every destination vertex: the final KDTree(k=1) transfer of predictions from
    the 24 000 sampled points back to the original mesh (inference_pipeline_tgn.py:137-140), as one grid kNN + gather."""
    from . import pointops
    L.require_cuda(src_xyz, dst_xyz)
    src = src_xyz.reshape(-1, 3).contiguous().float()
    dst = dst_xyz.reshape(-1, 3).contiguous().float()
    o = torch.tensor([src.shape[0]], dtype=torch.int32, device=src.device)
    no = torch.tensor([dst.shape[0]], dtype=torch.int32, device=src.device)
    idx, _ = pointops.knn_packed(1, src, dst, o, no)
    return src_labels.to(src.device).reshape(src.shape[0], -1)[idx[:, 0].long()].reshape(dst.shape[0], *src_labels.shape[1:])

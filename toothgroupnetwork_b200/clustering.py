"""The clustering step between the two stages of the reference's two-stage models, on the GPU.

``ops_utils.get_clustering_labels`` (ops_utils.py:86-144, called at grouping_network_module.py:65) runs
``sklearn.cluster.DBSCAN(eps=0.03, min_samples=30)`` on the offset-moved foreground points, splits suspiciously elongated
clusters with MeanShift and gives every noise point the majority label of its 10 nearest labelled points (a KDTree).  On a
trained network the DBSCAN alone costs 0.9 s per cloud on the host -- an order of magnitude more than the two networks.

Here ``dbscan`` is csrc/dbscan.cu (labels and core samples equal to scikit-learn's: same float64 predicate, same cluster
numbering, same border rule), the noise vote uses the float64 selection kernel of csrc/crop_knn.cu, and
``get_clustering_labels`` keeps the reference's steps and return value.  The elongation test is the reference's (PCA
eigenvalues of the core points); the MeanShift split, taken for at most three clusters and rarely, stays on scikit-learn.
``accelerate(ops_utils_module)`` swaps the function in an imported reference ``ops_utils``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _lib as L
from . import crops

_ws = {}


def _workspace(n: int, device) -> torch.Tensor:
    need = int(L.load().tgn_dbscan_bytes(int(n)))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def dbscan_device(points: torch.Tensor, eps: float = 0.03, min_samples: int = 30) -> Tuple[torch.Tensor, torch.Tensor]:
    """points (n,3) float32 CUDA -> (labels (n) int32, core (n) uint8), both on the device; equal to
    ``sklearn.cluster.DBSCAN(eps, min_samples).fit(points)``'s ``labels_`` / ``core_sample_indices_``."""
    L.require_cuda(points)
    if points.dim() != 2 or points.shape[1] != 3 or points.dtype != torch.float32:
        raise L.TgnError(f"dbscan expects (n,3) float32 points, got {tuple(points.shape)} {points.dtype}")
    points = points.contiguous()
    n = points.shape[0]
    labels = torch.empty(n, dtype=torch.int32, device=points.device)
    core = torch.empty(n, dtype=torch.uint8, device=points.device)
    if n:
        ws = _workspace(n, points.device)
        L.call("tgn_dbscan", n, L.ptr(points), float(eps), int(min_samples), L.ptr(ws), L.ptr(labels), L.ptr(core), None, L.stream_ptr())
    return labels, core


def _as_float32(points: np.ndarray) -> np.ndarray:
    """The kernel takes float32 coordinates (what ops_utils.get_clustering_labels is handed: float32 points + float32 offsets) and
    evaluates distances in float64 like scikit-learn.  Wider input is accepted when it holds float32 values; anything else would be
    clustered on rounded coordinates and could differ from scikit-learn, so it is refused."""
    points = np.asarray(points)
    if points.dtype == np.float32:
        return np.ascontiguousarray(points)
    down = np.ascontiguousarray(points, dtype=np.float32)
    if not np.array_equal(down.astype(points.dtype), points):
        raise L.TgnError(f"dbscan: {points.dtype} coordinates that are not float32 values are not supported")
    return down


def dbscan(points, eps: float = 0.03, min_samples: int = 30) -> Tuple[np.ndarray, np.ndarray]:
    """numpy or tensor (n,3) -> (labels_ int64 (n), core_sample_indices_ int64) like the attributes of a fitted sklearn DBSCAN."""
    pts = torch.as_tensor(_as_float32(points)) if isinstance(points, np.ndarray) else points.float()
    if not pts.is_cuda and not torch.cuda.is_available():
        raise L.TgnError("dbscan needs a CUDA device (there is no CPU path)")
    pts = pts if pts.is_cuda else pts.cuda()
    labels, core = dbscan_device(pts, eps, min_samples)
    both = torch.stack([labels, core.int()]).cpu().numpy()             # one device -> host copy
    return both[0].astype(np.int64), np.flatnonzero(both[1]).astype(np.int64)


class DBSCAN:
    """The slice of ``sklearn.cluster.DBSCAN`` the reference uses -- ``DBSCAN(eps=..., min_samples=...).fit(X, y)`` then
    ``.labels_`` / ``.core_sample_indices_`` / ``.components_`` -- on csrc/dbscan.cu: ops_utils.py:28,98, tsegnet.py:59,
    inference_pipeline_tsegnet.py:39.  Euclidean metric only, which is all the reference asks for."""

    def __init__(self, eps: float = 0.5, *, min_samples: int = 5, metric: str = "euclidean", **unsupported):
        if metric != "euclidean" or any(v is not None for v in unsupported.values()):
            raise L.TgnError(f"DBSCAN: only the Euclidean metric with default options is implemented (got metric={metric!r}, {unsupported})")
        self.eps, self.min_samples = eps, min_samples

    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise L.TgnError("DBSCAN: sample_weight is not implemented")
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[1] != 3:
            raise L.TgnError(f"DBSCAN: expected (n, 3) points, got {X.shape}")
        self.labels_, self.core_sample_indices_ = dbscan(X, self.eps, self.min_samples)
        self.components_ = X[self.core_sample_indices_].copy()
        self.n_features_in_ = 3
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, y, sample_weight).labels_


def _explained_variance(points: np.ndarray) -> np.ndarray:
    """sklearn PCA(n_components=3).fit(points).explained_variance_ (ops_utils.get_eg_values :51-56): eigenvalues of the
    sample covariance, descending."""
    if points.shape[0] < 3:
        return np.array([0, 0, 0])
    x = np.asarray(points, dtype=np.float64)
    x = x - x.mean(0)
    ev = np.linalg.eigvalsh(x.T @ x / (x.shape[0] - 1))[::-1]
    return np.maximum(ev, 0.0)


def _majority(rows: np.ndarray) -> np.ndarray:
    """per row: the most frequent value, the smallest one among equals (np.unique + argmax, ops_utils.py:139-141)."""
    values, inv = np.unique(rows, return_inverse=True)
    inv = inv.reshape(rows.shape)
    counts = np.zeros((rows.shape[0], values.shape[0]), dtype=np.int32)
    np.add.at(counts, (np.arange(rows.shape[0])[:, None], inv), 1)
    return values[np.argmax(counts, axis=1)]


def get_clustering_labels(moved_points, labels):
    """Drop-in for ops_utils.get_clustering_labels (:86-144): moved_points (N,3), labels (N,) or (N,1) semantic classes ->
    cluster label of every foreground point (labels != 0), int64."""
    if not torch.cuda.is_available():
        raise L.TgnError("get_clustering_labels needs a CUDA device (there is no CPU path)")
    moved_points = np.asarray(moved_points)
    cond = np.asarray(labels) != 0
    fg = moved_points[cond, :]
    fg_dev = torch.as_tensor(_as_float32(fg)).cuda()
    lab_dev, core_dev = dbscan_device(fg_dev, 0.03, 30)
    both = torch.stack([lab_dev, core_dev.int()]).cpu().numpy()
    clustering_labels = both[0].astype(np.int64)
    core_mask = both[1].astype(bool)
    dbscan_labels = clustering_labels.copy()

    uniq = [int(u) for u in np.unique(dbscan_labels) if u != -1]
    eg_values = np.array([_explained_variance(fg[core_mask & (dbscan_labels == u), :3]) for u in uniq])
    eg_first = eg_values[:, 0]
    order = np.argsort(-eg_first)
    eg_first = eg_first[order]
    suspicious = []
    for i in range(3):
        with np.errstate(invalid="ignore", divide="ignore"):
            if eg_first[i] / eg_first[3:].mean() > 8:
                suspicious.append(order[i])
    for idx, which in enumerate(suspicious):
        from sklearn.cluster import MeanShift                      # host-side, rare: at most three elongated clusters
        # the reference indexes its list of clusters by position and then rewrites the label equal to that position (:128-132)
        members = fg[dbscan_labels == uniq[which], :3].astype(np.float64)
        split = MeanShift(bandwidth=0.07).fit(members)
        clustering_labels[clustering_labels == which] = split.labels_ + 100 * (idx + 1)

    noise = clustering_labels == -1
    if noise.any():          # (with no noise at all the reference raises ValueError from KDTree.query on an empty array, :135)
        keep = ~noise
        if int(keep.sum()) < 10:
            raise ValueError("k must be less than or equal to the number of training points")
        near = crops.nearest_neighbor_crops(fg_dev[torch.as_tensor(keep).cuda()].unsqueeze(0), fg_dev[torch.as_tensor(noise).cuda()].unsqueeze(0), 10)[0]
        clustering_labels[noise] = _majority(clustering_labels[keep][near.cpu().numpy()])
    return clustering_labels


def accelerate(ops_utils_module) -> None:
    """Swap ops_utils.get_clustering_labels of an imported reference module for the device version, and the ``DBSCAN`` name the module
    imported from scikit-learn (``clustering_points(method="dbscan")``, ops_utils.py:28) for the class above."""
    ops_utils_module.get_clustering_labels = get_clustering_labels
    accelerate_dbscan(ops_utils_module)


def accelerate_dbscan(module) -> None:
    """Point a module's ``DBSCAN`` name (``from sklearn.cluster import DBSCAN`` in ops_utils.py, models/modules/tsegnet.py,
    inference_pipelines/inference_pipeline_tsegnet.py) at the device implementation."""
    if hasattr(module, "DBSCAN"):
        module.DBSCAN = DBSCAN

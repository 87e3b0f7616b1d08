"""DBSCAN(eps=0.03, min_samples=30) and the whole get_clustering_labels step: scikit-learn on the host (what the reference runs,
ops_utils.py:86-144) beside csrc/dbscan.cu, on offset-moved 24 000-point arches from 'untrained' to 'converged'."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sklearn.cluster import DBSCAN  # noqa: E402

from toothgroupnetwork_b200 import clouds, clustering  # noqa: E402


def moved_cloud(n, seed, pull, jitter):
    xyz, _, label = clouds.dental_arch(n, seed)
    xyz, label = xyz.numpy(), label.numpy().astype(np.int64)
    label = np.where(label < 0, 0, label)
    cent = np.stack([xyz[label == c].mean(0) if (label == c).any() else np.zeros(3, np.float32) for c in range(int(label.max()) + 1)])
    rng = np.random.default_rng(seed)
    return (xyz + pull * (cent[label] - xyz) + rng.normal(0, jitter, xyz.shape)).astype(np.float32), label


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    return out, float(np.median(ts))


def main():
    rows = []
    for pull, jitter in ((0.0, 0.0), (0.5, 0.004), (0.9, 0.004), (1.0, 0.0005)):
        pts, label = moved_cloud(24000, 0, pull, jitter)
        fg = pts[label != 0]
        ref, ms_ref = best_of(lambda: DBSCAN(eps=0.03, min_samples=30).fit(fg), 2)
        clustering.dbscan(fg)
        (lab, core), ms_ours = best_of(lambda: clustering.dbscan(fg), 5)
        dev = torch.as_tensor(fg).cuda()
        _, ms_dev = best_of(lambda: clustering.dbscan_device(dev), 5)
        (g, _), ms_full = best_of(lambda: (clustering.get_clustering_labels(pts, label), 0), 3)
        rows.append({"pull": pull, "jitter": jitter, "points": int(len(fg)), "clusters": int(ref.labels_.max() + 1), "noise": int((ref.labels_ == -1).sum()),
                     "core": int(len(ref.core_sample_indices_)), "sklearn_ms": ms_ref, "device_ms_numpy_in_out": ms_ours, "device_ms_resident": ms_dev,
                     "get_clustering_labels_ms": ms_full, "labels_equal": bool(np.array_equal(lab, ref.labels_)),
                     "core_equal": bool(np.array_equal(core, ref.core_sample_indices_))})
        print(json.dumps(rows[-1]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_cluster_bench.json"), "w") as f:
        json.dump({"host_threads": os.cpu_count(), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

"""GPU crop extraction (csrc/crop_knn.cu, toothgroupnetwork_b200/crops.py) against what the reference runs on the host:
sklearn KDTree.query(k=3072) (ops_utils.py:146-161) and a float64 brute force; label transfer against KDTree(k=1)."""
import numpy as np
import pytest
import torch

from toothgroupnetwork_b200 import clouds, crops

pytestmark = pytest.mark.gpu


def brute(xyz: np.ndarray, centres: np.ndarray, k: int) -> np.ndarray:
    d = ((xyz[None, :, :].astype(np.float64) - centres[:, None, :].astype(np.float64)) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    order = np.lexsort((np.broadcast_to(np.arange(xyz.shape[0]), d.shape), d), axis=-1)
    return order[:, :k]


def tooth_centroids(n=24000, seed=0):
    xyz, _, label = clouds.dental_arch(n, seed)
    cents = np.stack([xyz[label == t].numpy().mean(0) for t in range(16)]).astype(np.float32)
    return xyz, cents


@pytest.mark.parametrize("k", [3072, 4096, 1, 100])
def test_crop_knn_matches_float64_bruteforce(k):
    xyz, cents = tooth_centroids()
    got = crops.nearest_neighbor_crops(xyz[None].cuda(), torch.from_numpy(cents)[None].cuda(), k)[0].cpu().numpy()
    assert np.array_equal(got, brute(xyz.numpy(), cents, k))


def test_crop_knn_matches_sklearn_kdtree_and_dropin_signature():
    from sklearn.neighbors import KDTree
    B = 2
    clouds_np, cents = [], []
    for b in range(B):
        xyz, c = tooth_centroids(24000, 10 + b)
        clouds_np.append(xyz.numpy())
        cents.append(list(c[: 14 + b]))                       # ragged numbers of centres per cloud, as lists of arrays
    org = np.stack(clouds_np)
    want = []
    for b in range(B):                                        # ops_utils.py:154-160
        tree = KDTree(org[b], leaf_size=2)
        want.append(tree.query(cents[b], k=3072, return_distance=False))
    got = crops.get_nearest_neighbor_idx(org, cents, 3072)
    assert isinstance(got, list) and all(isinstance(g, np.ndarray) for g in got)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    # tensor in -> device tensors out, and the gather of ops_utils.get_indexed_features (:198-218)
    feats = torch.randn(B, 6, 24000, generator=torch.Generator().manual_seed(1)).cuda()
    got_t = crops.get_nearest_neighbor_idx(torch.from_numpy(org).cuda(), cents, 3072)
    assert all(t.is_cuda for t in got_t)
    cropped = crops.get_indexed_features(feats, got_t)
    ref = torch.stack([feats[b][:, want[b][q]] for b in range(B) for q in range(len(want[b]))], 0)
    assert torch.equal(cropped, ref)


def test_crop_knn_duplicates_short_cloud_and_far_centre():
    xyz = clouds.with_duplicates(clouds.cube(3000, 2), 3)                 # exact float64 ties: lower index first
    cents = np.array([[0.1, 0.2, 0.3], [5.0, 5.0, 5.0], xyz[17].numpy()], np.float32)
    got = crops.nearest_neighbor_crops(xyz[None].cuda(), torch.from_numpy(cents)[None].cuda(), 2048)[0].cpu().numpy()
    assert np.array_equal(got, brute(xyz.numpy(), cents, 2048))
    small = clouds.cube(100, 5)
    got = crops.nearest_neighbor_crops(small[None].cuda(), torch.from_numpy(cents[:1])[None].cuda(), 128)[0, 0].cpu().numpy()
    assert np.array_equal(got[:100], brute(small.numpy(), cents[:1], 100)[0]) and (got[100:] == 0).all()


def test_nearest_label_transfer_matches_kdtree():
    from sklearn.neighbors import KDTree
    src, _, lab = clouds.dental_arch(24000, 3)
    dst = clouds.dental_arch(100000, 4)[0]
    want = lab.numpy()[KDTree(src.numpy(), leaf_size=2).query(dst.numpy(), k=1, return_distance=False)[:, 0]]
    got = crops.nearest_label_transfer(src.cuda(), lab.cuda(), dst.cuda()).cpu().numpy()
    # float32 (kernel) vs float64 (KDTree) nearest neighbour can differ only where two candidates are equidistant to ~1e-7
    assert (got != want).mean() < 1e-4

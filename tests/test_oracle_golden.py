"""The oracle (CPU restatement) against fixtures produced by the reference's own python code
(tests/golden/make_ref_torch_golden.py).  Bit-exact for indices and expanded distances;
MLP outputs within 1e-4 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _layers(fix, n, prefix=""):
    out = []
    for i in range(n):
        t = lambda k: torch.from_numpy(fix[f"{prefix}{k}{i}"])
        out.append(oracle.MlpParams(t("w"), t("b"), t("gamma"), t("beta"), t("mean"), t("var")))
    return out


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def test_square_distance_bit_exact(golden_dir):
    fix = _load(golden_dir, "ref_torch_ball.npz")
    xyz = fix["xyz"]
    new = xyz[fix["sel"]]
    sd = oracle.square_distance(new[None], xyz[None])[0]
    assert np.array_equal(sd[:8].view(np.uint32), fix["sqdist_rows"].view(np.uint32))


@pytest.mark.parametrize("r,k", [(0.025, 32), (0.05, 64), (0.1, 32), (0.2, 16)])
def test_query_ball_point_exact(golden_dir, r, k):
    fix = _load(golden_dir, "ref_torch_ball.npz")
    xyz = fix["xyz"]
    new = xyz[fix["sel"]]
    got = oracle.query_ball_point(r, k, xyz[None], new[None])[0]
    assert np.array_equal(got, fix[f"ball_r{r}_k{k}"].astype(np.int64))


def test_query_ball_point_empty_ball_sentinel(golden_dir):
    fix = _load(golden_dir, "ref_torch_ball.npz")
    xyz = fix["xyz"]
    got = oracle.query_ball_point(0.1, 8, xyz[None], np.array([[[5.0, 5.0, 5.0]]], np.float32))[0]
    assert np.array_equal(got, fix["ball_far"].astype(np.int64))
    assert (got == xyz.shape[0]).all()


def test_three_nn_matches_sorted_matrix(golden_dir):
    fix = _load(golden_dir, "ref_torch_fp.npz")
    x1 = fix["xyz1"]
    x2 = x1[fix["fps"]]
    d, i = oracle.three_nn(x1[None], x2[None])
    assert np.array_equal(d[0].view(np.uint32), fix["nn3_d"].view(np.uint32))
    # indices may legitimately differ only where distances tie exactly
    ties = (fix["nn3_d"][:, 0] == fix["nn3_d"][:, 1]) | (fix["nn3_d"][:, 1] == fix["nn3_d"][:, 2])
    assert np.array_equal(i[0][~ties], fix["nn3_idx"][~ties].astype(np.int64))


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_feature_propagation(golden_dir, mode):
    fix = _load(golden_dir, "ref_torch_fp.npz")
    x1 = torch.from_numpy(fix["xyz1"])
    x2 = x1[torch.from_numpy(fix["fps"]).long()]
    out = oracle.feature_propagation(x1.t()[None], x2.t()[None], torch.from_numpy(fix["points1"]),
                                     torch.from_numpy(fix["points2"]), _layers(fix, 2), train_bn=(mode == "train"))
    assert _rel(out.numpy(), fix[f"out_{mode}"]) < 1e-4


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_set_abstraction_ssg(golden_dir, mode):
    fix = _load(golden_dir, "ref_torch_sa.npz")
    feats = torch.from_numpy(fix["feats"])
    nx, npts = oracle.set_abstraction(feats[:, :3].contiguous(), feats, 128, 0.1, 32, _layers(fix, 3),
                                      train_bn=(mode == "train"))
    assert np.array_equal(nx.numpy(), fix[f"new_xyz_{mode}"])
    assert _rel(npts.numpy(), fix[f"new_points_{mode}"]) < 1e-4


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_set_abstraction_msg(golden_dir, mode):
    fix = _load(golden_dir, "ref_torch_msg.npz")
    feats = torch.from_numpy(fix["feats"])
    branches = [_layers(fix, 2, "br0_"), _layers(fix, 2, "br1_")]
    nx, npts = oracle.set_abstraction_msg(feats[:, :3].contiguous(), feats, 128, [0.05, 0.1], [16, 32], branches,
                                          train_bn=(mode == "train"))
    assert np.array_equal(nx.numpy(), fix[f"new_xyz_{mode}"])
    assert _rel(npts.numpy(), fix[f"new_points_{mode}"]) < 1e-4


def test_set_abstraction_msg_real_sa1_shape(golden_dir):
    """Oracle against the reference-Python fixture of the real pointnet_pp SA1 shape (9->128->128 twice)."""
    fix = _load(golden_dir, "ref_torch_msg128.npz")
    feats = torch.from_numpy(fix["feats"])
    branches = [_layers(fix, 2, "br0_"), _layers(fix, 2, "br1_")]
    nx, npts = oracle.set_abstraction_msg(feats[:, :3].contiguous(), feats, 128, [0.05, 0.1], [32, 64], branches)
    assert np.array_equal(nx.numpy(), fix["new_xyz_eval"])
    assert _rel(npts.numpy(), fix["new_points_eval"]) < 1e-4


def test_set_abstraction_group_all(golden_dir):
    fix = _load(golden_dir, "ref_torch_groupall.npz")
    feats = torch.from_numpy(fix["feats"])
    nx, npts = oracle.set_abstraction(feats[:, :3].contiguous(), feats, None, None, None, _layers(fix, 2),
                                      group_all=True)
    assert np.array_equal(nx.numpy(), fix["new_xyz_eval"])
    assert _rel(npts.numpy(), fix["new_points_eval"]) < 1e-4


def test_fps_matches_first_index_argmax_on_tie_free_cloud():
    """On a tie-free random cloud any correct fp32 FPS gives the same indices
    (SURVEY.md 7.1): cross-check the block-emulating oracle against a plain numpy FPS that uses
    the same fma-ordered distance."""
    from toothgroupnetwork_b200 import clouds
    xyz = clouds.cube(3000, seed=1).numpy()
    got = oracle.furthestsampling(xyz, [3000], [200])
    d = np.full(3000, 1e10, np.float32)
    cur, ref = 0, [0]
    for _ in range(199):
        diff = (xyz - xyz[cur]).astype(np.float32)
        # fma(dz,dz,fma(dx,dx,dy*dy)) emulated in float64 then rounded: exact for this check
        t = (diff[:, 1].astype(np.float64) ** 2).astype(np.float32)
        t = (diff[:, 0].astype(np.float64) ** 2 + t.astype(np.float64)).astype(np.float32)
        t = (diff[:, 2].astype(np.float64) ** 2 + t.astype(np.float64)).astype(np.float32)
        d = np.minimum(d, t)
        cur = int(np.argmax(d))
        ref.append(cur)
    assert np.array_equal(got, np.array(ref, np.int32))


def test_fps_ragged_batch_and_block_size():
    from toothgroupnetwork_b200 import clouds
    a, b = clouds.cube(700, 2).numpy(), clouds.cube(93, 3).numpy()
    xyz = np.concatenate([a, b])
    idx = oracle.furthestsampling(xyz, [700, 793], [64, 76])
    assert idx[0] == 0 and idx[64] == 700
    assert (idx[:64] < 700).all() and (idx[64:] >= 700).all()
    assert len(set(idx.tolist())) == 76
    assert oracle.opt_n_threads(24000) == 1024 and oracle.opt_n_threads(93) == 64 and oracle.opt_n_threads(1) == 1


def test_knn_sorted_and_padding():
    from toothgroupnetwork_b200 import clouds
    xyz = clouds.cube(500, 4).numpy()
    idx, dist, d2 = oracle.knnquery(8, xyz, xyz[:50], [500], [50])
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (idx[:, 0] == np.arange(50)).all()
    # segment smaller than k: trailing slots keep (start, 1e10)  (knnquery_cuda_kernel.cu:88-91)
    idx, dist, d2 = oracle.knnquery(8, xyz[:5], xyz[:2], [5], [2])
    assert (idx[:, 5:] == 0).all() and (d2[:, 5:] == np.float32(1e10)).all()

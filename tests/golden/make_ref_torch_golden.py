"""Generate tests/golden/ref_torch_*.npz by IMPORTING THE PYTHON REFERENCE from /root/reference
and running its own functions on CPU (this container only; the GPU box has no /root/reference).

    python tests/golden/make_ref_torch_golden.py

What is pinned (reference file: external_libs/pointnet2_utils/pointnet2_utils.py):
  * square_distance (:20-41), query_ball_point (:120-144), index_points (:44-61) -- pure torch,
    CPU-runnable unmodified;
  * PointNetFeaturePropagation.forward (:313-352) -- CPU-runnable unmodified;
  * PointNetSetAbstraction / PointNetSetAbstractionMsg .forward (:213-239, :261-299).  These
    hard-code ``.cuda()`` inside farthest_point_sample (:88-96) and call the CUDA-only
    ``pointops_cuda.furthestsampling_cuda``; to run them here the script (a) registers a stub
    ``pointops_cuda`` module whose FPS is oracle.furthestsampling (itself pinned against the
    verbatim reference kernel by ref_cuda_*.npz) and (b) makes ``Tensor.cuda`` /
    ``torch.cuda.*Tensor`` CPU no-ops for the duration of the call.  Everything else executed
    is the reference's own code.

Precision flags: CPU fp32 (no TF32 anywhere).  BatchNorm mode is recorded per fixture.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import oracle  # noqa: E402
from toothgroupnetwork_b200 import clouds  # noqa: E402


def install_stubs():
    stub = types.ModuleType("pointops_cuda")

    def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        out = oracle.furthestsampling(xyz.numpy(), offset.numpy(), new_offset.numpy())
        idx.copy_(torch.from_numpy(out))

    stub.furthestsampling_cuda = furthestsampling_cuda
    sys.modules["pointops_cuda"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.IntTensor = lambda *a: torch.zeros(*[int(x) for x in a], dtype=torch.int32)
    torch.cuda.FloatTensor = lambda *a: torch.zeros(*[int(x) for x in a], dtype=torch.float32)
    sys.path.insert(0, REF)


def dump_mlp(convs, bns):
    d = {}
    for i, (c, b) in enumerate(zip(convs, bns)):
        d[f"w{i}"] = c.weight.detach().reshape(c.weight.shape[0], -1).numpy()
        d[f"b{i}"] = c.bias.detach().numpy()
        d[f"gamma{i}"] = b.weight.detach().numpy()
        d[f"beta{i}"] = b.bias.detach().numpy()
        d[f"mean{i}"] = b.running_mean.detach().numpy().copy()
        d[f"var{i}"] = b.running_var.detach().numpy().copy()
    return d


def randomize_bn(bns, gen):
    for b in bns:
        with torch.no_grad():
            b.weight.copy_(torch.rand(b.weight.shape, generator=gen) + 0.5)
            b.bias.copy_(torch.randn(b.bias.shape, generator=gen) * 0.1)
            b.running_mean.copy_(torch.randn(b.running_mean.shape, generator=gen) * 0.1)
            b.running_var.copy_(torch.rand(b.running_var.shape, generator=gen) + 0.5)


def main():
    install_stubs()
    import warnings
    warnings.filterwarnings("ignore")
    from external_libs.pointnet2_utils import pointnet2_utils as ref

    # ---- square_distance / ball query / index_points on an arch cloud -------------------
    xyz, normal, _ = clouds.dental_arch(4096, seed=3)
    g = torch.Generator().manual_seed(11)
    sel = torch.randperm(4096, generator=g)[:160]
    new_xyz = xyz[sel]
    sd = ref.square_distance(new_xyz[None], xyz[None])[0]
    fix = {"xyz": xyz.numpy(), "sel": sel.numpy().astype(np.int32), "sqdist_rows": sd[:8].numpy()}
    for r, k in [(0.025, 32), (0.05, 64), (0.1, 32), (0.2, 16)]:
        gi = ref.query_ball_point(r, k, xyz[None], new_xyz[None])[0]
        fix[f"ball_r{r}_k{k}"] = gi.numpy().astype(np.int32)
    # a query far from the cloud -> the sentinel N row
    far = torch.tensor([[[5.0, 5.0, 5.0]]])
    fix["ball_far"] = ref.query_ball_point(0.1, 8, xyz[None], far)[0].numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "ref_torch_ball.npz"), **fix)

    # ---- feature propagation ---------------------------------------------------------------
    torch.manual_seed(0)
    N, S = 2048, 256
    x1, n1, _ = clouds.dental_arch(N, seed=5)
    fps = oracle.furthestsampling(x1.numpy(), np.array([N], np.int32), np.array([S], np.int32))
    x2 = x1[torch.from_numpy(fps).long()]
    gen = torch.Generator().manual_seed(21)
    p1 = torch.randn(1, 6, N, generator=gen)
    p2 = torch.randn(1, 24, S, generator=gen)
    fp = ref.PointNetFeaturePropagation(30, [32, 16])
    randomize_bn(fp.mlp_bns, gen)
    fix = {"xyz1": x1.numpy(), "fps": fps, "points1": p1.numpy(), "points2": p2.numpy()}
    fix.update(dump_mlp(fp.mlp_convs, fp.mlp_bns))
    fp.eval()
    with torch.no_grad():
        fix["out_eval"] = fp(x1.t()[None], x2.t()[None], p1, p2).numpy()
    fp.train()
    with torch.no_grad():
        fix["out_train"] = fp(x1.t()[None], x2.t()[None], p1, p2).numpy()
    # interpolation alone (no skip, no mlp): the 3-NN idx / weights of the reference body
    d = ref.square_distance(x1[None], x2[None])
    dd, ii = d.sort(dim=-1)
    fix["nn3_idx"] = ii[0, :, :3].numpy().astype(np.int32)
    fix["nn3_d"] = dd[0, :, :3].numpy()
    np.savez_compressed(os.path.join(OUT, "ref_torch_fp.npz"), **fix)

    # ---- set abstraction (SSG, BASELINE C2(i) shape at reduced N) and MSG -----------------
    torch.manual_seed(0)
    B, N, S = 2, 2048, 128
    feats = torch.cat([clouds.arch_features(N, seed=7), clouds.arch_features(N, seed=8)], 0)  # (B,6,N)
    gen = torch.Generator().manual_seed(31)
    sa = ref.PointNetSetAbstraction(S, 0.1, 32, 9, [32, 32, 64], False)
    randomize_bn(sa.mlp_bns, gen)
    fix = {"feats": feats.numpy()}
    fix.update(dump_mlp(sa.mlp_convs, sa.mlp_bns))
    for mode in ("eval", "train"):
        sa.train(mode == "train")
        with torch.no_grad():
            nx, npts = sa(feats[:, :3].contiguous(), feats)
        fix[f"new_xyz_{mode}"] = nx.numpy()
        fix[f"new_points_{mode}"] = npts.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_torch_sa.npz"), **fix)

    msg = ref.PointNetSetAbstractionMsg(S, [0.05, 0.1], [16, 32], 6, [[16, 32], [32, 48]])
    fix = {"feats": feats.numpy()}
    for bi in range(2):
        randomize_bn(msg.bn_blocks[bi], gen)
        for k, v in dump_mlp(msg.conv_blocks[bi], msg.bn_blocks[bi]).items():
            fix[f"br{bi}_{k}"] = v
    for mode in ("eval", "train"):
        msg.train(mode == "train")
        with torch.no_grad():
            nx, npts = msg(feats[:, :3].contiguous(), feats)
        fix[f"new_xyz_{mode}"] = nx.numpy()
        fix[f"new_points_{mode}"] = npts.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_torch_msg.npz"), **fix)

    # group_all variant (tsg_seg_module.py:27 flatten_sa)
    ga = ref.PointNetSetAbstraction(None, None, None, 6 + 3, [16, 32], True)
    randomize_bn(ga.mlp_bns, gen)
    fix = {"feats": feats[:, :, :256].numpy()}
    fix.update(dump_mlp(ga.mlp_convs, ga.mlp_bns))
    ga.eval()
    with torch.no_grad():
        nx, npts = ga(feats[:, :3, :256].contiguous(), feats[:, :, :256].contiguous())
    fix["new_xyz_eval"], fix["new_points_eval"] = nx.numpy(), npts.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_torch_groupall.npz"), **fix)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

"""Generate tests/golden/ref_torch_msg128.npz: the REAL pointnet_pp SA1 of the reference
(models/modules/pointnet_pp.py:13, PointNetSetAbstractionMsg(1024, [0.025, 0.05], [32, 64], 6,
[[128, 128], [128, 128]])) at reduced size (N = 4096, S = 128; the radii are doubled so that the
balls of the sparser cloud are populated), run by IMPORTING THE PYTHON REFERENCE on CPU exactly
like make_ref_torch_golden.py does (same stubs, same helpers).  Eval-mode BatchNorm with
randomised running statistics.

    python tests/golden/make_ref_torch_msg128_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_ref_torch_golden as g  # noqa: E402


def main():
    g.install_stubs()
    warnings.filterwarnings("ignore")
    from external_libs.pointnet2_utils import pointnet2_utils as ref

    torch.manual_seed(0)
    N, S = 4096, 128
    feats = torch.cat([g.clouds.arch_features(N, seed=17), g.clouds.arch_features(N, seed=18)], 0)   # (2,6,N)
    gen = torch.Generator().manual_seed(77)
    msg = ref.PointNetSetAbstractionMsg(S, [0.05, 0.1], [32, 64], 6, [[128, 128], [128, 128]])
    fix = {"feats": feats.numpy()}
    for bi in range(2):
        g.randomize_bn(msg.bn_blocks[bi], gen)
        for k, v in g.dump_mlp(msg.conv_blocks[bi], msg.bn_blocks[bi]).items():
            fix[f"br{bi}_{k}"] = v
    msg.eval()
    with torch.no_grad():
        nx, npts = msg(feats[:, :3].contiguous(), feats)
    fix["new_xyz_eval"], fix["new_points_eval"] = nx.numpy(), npts.numpy()
    out = os.path.join(g.OUT, "ref_torch_msg128.npz")
    np.savez_compressed(out, **fix)
    print(out, os.path.getsize(out), "bytes; new_points", npts.shape, float(npts.abs().max()))


if __name__ == "__main__":
    main()

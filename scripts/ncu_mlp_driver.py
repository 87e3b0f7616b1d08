"""One launch of each single-kernel set-abstraction MLP engine on the bench shape (for ncu)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from toothgroupnetwork_b200 import pointnet2_utils as pn2
B = 592
sa = bench.build_module("cuda")
feats = bench.make_clouds(0, B).cuda()
xyz = feats[:, :3].contiguous()
xyz_t, feats_t = pn2.transpose_last2(xyz), pn2.transpose_last2(feats)
fps = pn2._fps_batched(xyz_t, bench.NPOINT)
new_xyz_t = pn2._take_rows(xyz_t.view(-1, 3), fps).view(B, bench.NPOINT, 3)
gidx = pn2._ball_query(bench.RADIUS, bench.NSAMPLE, xyz_t, new_xyz_t, False, None, 1)
folded = sa._folded.update(sa.mlp_convs, sa.mlp_bns)
out = torch.empty((B, 64, bench.NPOINT), device="cuda")
for eng in (4, 2):
    for _ in range(2):
        pn2.sa_group_mlp_max(xyz_t, feats_t, new_xyz_t, gidx, True, folded, out, 0, eng)
torch.cuda.synchronize()

"""One FPS launch per schedule at B=1 (24000 -> 1024), for ncu source-level stall sampling."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import op_bench
from toothgroupnetwork_b200 import pointops
B, M = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 1024
feats = op_bench.arch_batch(B, 24000)
xyz = feats[:, :3].permute(0, 2, 1).contiguous().view(-1, 3)
off = (torch.arange(1, B + 1, dtype=torch.int32) * 24000).cuda()
noff = (torch.arange(1, B + 1, dtype=torch.int32) * M).cuda()
for mode in (-26, -56):
    for _ in range(2):
        pointops.fps_packed(xyz, off, noff, 24000, B * M, mode)
torch.cuda.synchronize()

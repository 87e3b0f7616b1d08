"""Tile vs grid ball query over radii / sizes (experiments; prints ms per call)."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from toothgroupnetwork_b200 import clouds, pointnet2_utils as pn2

def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
for N, S in ((24000, 1024), (16384, 1024), (65536, 2048)):
    base = [clouds.dental_arch(N, s)[0] for s in range(4)]
    xyz = torch.stack([base[i % 4] for i in range(B if N < 60000 else B // 4)]).cuda().contiguous()
    sel = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:S].cuda()
    new = xyz[:, sel].contiguous()
    for K in (16, 32, 64):
        for r in (0.025, 0.05, 0.075, 0.1, 0.15, 0.2):
            t_tile = timeit(lambda: pn2._ball_query(r, K, xyz, new, False, pn2.BALL_TILE))
            t_grid = timeit(lambda: pn2._ball_query(r, K, xyz, new, False, pn2.BALL_GRID))
            t_auto = timeit(lambda: pn2._ball_query(r, K, xyz, new, False, pn2.BALL_AUTO))
            print(f"N={N} S={S} K={K} r={r}: tile {t_tile:.3f} grid {t_grid:.3f} auto {t_auto:.3f} ms", flush=True)

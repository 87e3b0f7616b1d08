"""ball query + group-MLP sweep, N in {4096,16384,24000,65536}, S = N/4, K in {16,32,64}, r = 0.1,
MLP 9->[32,32,64] (eval BN), device-resident; also FPS N->S.  Prints one markdown row per configuration."""
import sys, torch
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

points_budget = 14_000_000        # points per batch (about half of the bench batch)
print("| N | S | K | clouds | FPS ms | ball ms | MLP ms | sampled pts/s (ball+MLP) | GFLOP/s MLP | grouped GB/s MLP |")
print("|---|---|---|---|---|---|---|---|---|---|")
for N in (4096, 16384, 24000, 65536):
    S = N // 4
    B = max(4, points_budget // N)
    base = [clouds.arch_features(N, s) for s in range(4)]
    feats = torch.cat([base[i % 4] for i in range(B)], 0).cuda().contiguous()
    xyz = feats[:, :3].contiguous()
    xyz_t, feats_t = pn2.transpose_last2(xyz), pn2.transpose_last2(feats)
    fps = pn2._fps_batched(xyz_t, S)
    t_fps = timeit(lambda: pn2._fps_batched(xyz_t, S), 3)
    new_xyz_t = pn2._take_rows(xyz_t.view(-1, 3), fps).view(B, S, 3)
    for K in (16, 32, 64):
        sa = pn2.PointNetSetAbstraction(S, 0.1, K, 9, [32, 32, 64], False).cuda().eval()
        folded = sa._folded.update(sa.mlp_convs, sa.mlp_bns)
        gidx = pn2._ball_query(0.1, K, xyz_t, new_xyz_t, False)
        out = torch.empty((B, 64, S), device="cuda")
        t_ball = timeit(lambda: pn2._ball_query(0.1, K, xyz_t, new_xyz_t, False))
        t_mlp = timeit(lambda: pn2.sa_group_mlp_max(xyz_t, feats_t, new_xyz_t, gidx, True, folded, out, 0))
        rows = B * S * K
        flop = 2.0 * rows * (9 * 32 + 32 * 32 + 32 * 64)
        print(f"| {N} | {S} | {K} | {B} | {t_fps:.2f} | {t_ball:.3f} | {t_mlp:.3f} | {B * S / ((t_ball + t_mlp) * 1e-3):.3e} | "
              f"{flop / (t_mlp * 1e-3) / 1e9:.0f} | {rows * 9 * 4 / (t_mlp * 1e-3) / 1e9:.0f} |", flush=True)

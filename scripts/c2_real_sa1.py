"""the real pointnet_pp SA1, PointNetSetAbstractionMsg(1024,[0.025,0.05],[32,64],6,[[128,128],[128,128]]),
eval BN, batch of 24k-point clouds; stage times and which engine answered."""
import sys, torch
sys.path.insert(0, ".")
from toothgroupnetwork_b200 import clouds, pointnet2_utils as pn2, _lib as L

def timeit(f, n=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 148, 24000
base = [clouds.arch_features(N, s) for s in range(4)]
feats = torch.cat([base[i % 4] for i in range(B)], 0).cuda().contiguous()
xyz = feats[:, :3].contiguous()
msg = pn2.PointNetSetAbstractionMsg(1024, [0.025, 0.05], [32, 64], 6, [[128, 128], [128, 128]]).cuda().eval()
with torch.no_grad():
    t = timeit(lambda: msg(xyz, feats))
    print(f"MSG SA1 forward, {B} clouds: {t:.2f} ms  -> {B * 1024 / (t * 1e-3):.3e} sampled points/s")
    xyz_t = pn2.transpose_last2(xyz)
    fps = pn2._fps_batched(xyz_t, 1024)
    new_xyz_t = pn2._take_rows(xyz_t.view(-1, 3), fps).view(B, 1024, 3)
    print(f"  FPS {timeit(lambda: pn2._fps_batched(xyz_t, 1024)):.2f} ms")
    for r, K in ((0.025, 32), (0.05, 64)):
        for path, name in ((pn2.BALL_AUTO, "auto"), (pn2.BALL_TILE, "tile"), (pn2.BALL_GRID, "grid")):
            print(f"  ball r={r} K={K} {name}: {timeit(lambda: pn2._ball_query(r, K, xyz_t, new_xyz_t, False, path)):.2f} ms")
    ms = pn2.PointNetSetAbstraction(1024, 0.05, 64, 9, [128, 128], False).cuda().eval()
    for eng, name in ((pn2.ENGINE_AUTO, "auto"), (pn2.ENGINE_FP32, "fp32"), (pn2.ENGINE_TCW, "tcw")):
        pn2.set_sa_engine(eng)
        try:
            print(f"  SSG r=0.05 K=64 9->[128,128] engine {name}: {timeit(lambda: ms(xyz, feats)):.2f} ms (incl. FPS + ball)")
        except Exception as e:
            print("  engine", name, "failed:", e)
    pn2.set_sa_engine(pn2.ENGINE_AUTO)

import sys, os, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from toothgroupnetwork_b200 import pointnet2_utils as pn2
from test_gpu_pointnet2 import load, layers_from, fill_module, rel_err
pn2.set_reference_device("cpu")
fix = load("tests/golden", "ref_torch_msg128.npz")
feats = torch.from_numpy(fix["feats"]).cuda()
print("feats", feats.shape)
msg = pn2.PointNetSetAbstractionMsg(128, [0.05, 0.1], [32, 64], 6, [[128, 128], [128, 128]]).cuda().eval()
for bi in range(2):
    fill_module(msg.conv_blocks[bi], msg.bn_blocks[bi], layers_from(fix, 2, f"br{bi}_"))
want = fix["new_points_eval"]
outs = {}
for eng in (pn2.ENGINE_AUTO, pn2.ENGINE_FP32, pn2.ENGINE_TCW):
    pn2.set_sa_engine(eng)
    with torch.no_grad():
        nx, npts = msg(feats[:, :3].contiguous(), feats)
    pn2.set_sa_engine(pn2.ENGINE_AUTO)
    o = npts.cpu().numpy(); outs[eng] = o
    print("engine", eng, "all %.3e  branch0 %.3e  branch1 %.3e" % (rel_err(o, want), rel_err(o[:, :128], want[:, :128]), rel_err(o[:, 128:], want[:, 128:])))
d = np.abs(outs[0] - want)
idx = np.unravel_index(d.argmax(), d.shape); print("worst at", idx, outs[0][idx], want[idx], "frac of elements off by >1e-3 abs:", float((d > 1e-3).mean()))
bad_s = np.unique(np.where(d > 1e-3)[2]); print("bad s columns", bad_s[:40], len(bad_s))
bad_c = np.unique(np.where(d > 1e-3)[1]); print("bad channels", bad_c[:40], len(bad_c))

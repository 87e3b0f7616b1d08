import sys, torch
sys.path.insert(0, ".")
from toothgroupnetwork_b200 import clouds, pointnet2_utils as pn2
torch.backends.cudnn.allow_tf32 = False
feats = clouds.arch_features(6000, 1).cuda()
xyz = feats[:, :3].contiguous()
def rel(a, b): return float(((a - b).abs() / b.abs().clamp(min=0.05 * float(b.abs().max()))).max())
for widths in ([128], [128, 128], [96, 128, 72]):
    for randomize in ("none", "meanvar", "gammabeta", "all"):
        torch.manual_seed(0)
        sa = pn2.PointNetSetAbstraction(256, 0.1, 32, 9, widths, False).cuda().eval()
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for bn in sa.mlp_bns:
                if randomize in ("meanvar", "all"):
                    bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.3)
                    bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
                if randomize in ("gammabeta", "all"):
                    bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) + 0.5)
                    bn.bias.copy_(torch.randn(bn.bias.shape, generator=g) * 0.1)
            got = sa(xyz, feats)[1]                        # auto: widths > 64 -> pw chain, eval mode
            pn2.set_sa_engine(pn2.ENGINE_FP32)
            want = sa(xyz, feats)[1]                       # exact-FMA fused engine (folded BN)
            pn2.set_sa_engine(pn2.ENGINE_AUTO)
            pn2.set_pw_enabled(False)
            # torch reference formulation through the unfused path needs grad; emulate: module under enable_grad
        pn2.set_pw_enabled(True)
        print(widths, randomize, "pw vs fp32 engine: %.3e" % rel(got, want))

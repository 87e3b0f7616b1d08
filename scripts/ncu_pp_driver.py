"""PointPpFirstModule.forward (the reference's models/modules/pointnet_pp.py through dropin) in train() mode under no_grad,
B = 1, 24k points: the launch list shows which kernels the set-abstraction / feature-propagation levels run."""
import sys, warnings, torch
sys.path.insert(0, ".")
warnings.filterwarnings("ignore")
from oracle import ref_models
from toothgroupnetwork_b200 import clouds
torch.backends.cudnn.allow_tf32 = False
w = ref_models.World("b200")
with w:
    torch.manual_seed(0)
    model = w.mod("models.modules.pointnet_pp").PointPpFirstModule({}).cuda().train()
    feats = clouds.arch_features(24000, 0).cuda()
    with torch.no_grad():
        for _ in range(2):
            out = model([feats])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measured_forward")
        out = model([feats])
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
print("cls_pred", tuple(out["cls_pred"].shape))

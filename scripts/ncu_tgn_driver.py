"""GroupingNetworkModule (tgnet_fps) forward + backward (or, with "nograd", the no-grad forward on the fused blocks) through dropin, for a launch list: which kernels the step spends its time in."""
import sys, warnings, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
warnings.filterwarnings("ignore")
from oracle import ref_models
import model_parity as mp
w = ref_models.World(sys.argv[1] if len(sys.argv) > 1 else "b200")
feats, labels = mp.make_inputs(24000)
with w:
    torch.manual_seed(0)
    module = w.mod("models.modules.grouping_network_module").GroupingNetworkModule({"model_parameter": dict(mp.TGN_PARAMS)}).cuda().train()
    nograd = len(sys.argv) > 2 and sys.argv[2] == "nograd"
    def step():
        if nograd:
            with torch.no_grad():
                module([feats, labels])
            return
        module.zero_grad(set_to_none=True)
        out = module([feats, labels])
        loss = out["cbl_loss_1"].sum() + out["cbl_loss_2"].sum() + (out["sem_1"] ** 2).mean() + (out["offset_1"] ** 2).mean() + (out["sem_2"] ** 2).mean()
        loss.backward()
    step(); step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured_step")
    step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")

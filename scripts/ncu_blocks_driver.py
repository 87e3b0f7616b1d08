"""One fused transformer layer (c = 32, K = 36, n = 24 000, batch-statistics BatchNorm) and one DBSCAN of a converged arch, for ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_blocks import Layer  # noqa: E402

from toothgroupnetwork_b200 import blocks_fused, clouds, clustering  # noqa: E402

torch.manual_seed(0)
n, c, K = 24000, 32, 36
layer = Layer(c, K).cuda().train()
p = clouds.dental_arch(n, 1)[0].cuda().contiguous()
x = torch.randn(n, c, device="cuda")
o = torch.tensor([n], dtype=torch.int32, device="cuda")
xyz, _, label = clouds.dental_arch(n, 0)
xyz, label = xyz.numpy(), np.where(label.numpy() < 0, 0, label.numpy()).astype(np.int64)
cent = np.stack([xyz[label == k].mean(0) if (label == k).any() else np.zeros(3, np.float32) for k in range(int(label.max()) + 1)])
fg = torch.as_tensor((xyz + 0.9 * (cent[label] - xyz) + np.random.default_rng(0).normal(0, 0.004, xyz.shape)).astype(np.float32)[label != 0]).cuda()
with torch.no_grad():
    for _ in range(2):
        blocks_fused.pt_layer_forward(layer, [p, x, o])
        clustering.dbscan_device(fg)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    blocks_fused.pt_layer_forward(layer, [p, x, o])
    clustering.dbscan_device(fg)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")

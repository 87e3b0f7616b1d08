#!/usr/bin/env python
"""Which association does torch.sum(x ** 2, -1) use on this GPU for a (.., 3) tensor?  (square_distance :39-40)"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from toothgroupnetwork_b200 import clouds
out = {}
for B, N in ((1, 24000), (16, 24000), (1, 1024), (16, 1024), (8, 3072), (1, 512), (2, 256)):
    x = torch.stack([clouds.arch_features(N, s)[0][:3].t().contiguous() for s in range(B)]).cuda()   # (B,N,3)
    s_t = torch.sum(x ** 2, -1).cpu().numpy().reshape(-1)
    xc = x.cpu().numpy().reshape(-1, 3)
    sq = (xc * xc).astype(np.float32)
    f = np.float32
    def fma(a, b, c): return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    cands = {
        "(x2+y2)+z2": (sq[:, 0] + sq[:, 1]) + sq[:, 2],
        "x2+(y2+z2)": sq[:, 0] + (sq[:, 1] + sq[:, 2]),
        "(x2+z2)+y2": (sq[:, 0] + sq[:, 2]) + sq[:, 1],
        "fma(z,z,fma(y,y,x*x))": fma(xc[:, 2], xc[:, 2], fma(xc[:, 1], xc[:, 1], sq[:, 0])),
        "fma(x,x,fma(y,y,z*z))": fma(xc[:, 0], xc[:, 0], fma(xc[:, 1], xc[:, 1], sq[:, 2])),
        "f64 sum": (sq.astype(np.float64).sum(1)).astype(np.float32),
    }
    out[f"B{B}_N{N}"] = {k: int((v.astype(np.float32).view(np.int32) != s_t.view(np.int32)).sum()) for k, v in cands.items()}
    # non-contiguous variant as the reference calls it: src is new_xyz (B,S,3) contiguous; dst = xyz (B,N,3) from permute?  both contiguous in our harness
    xp = x.permute(0, 2, 1).contiguous().permute(0, 2, 1)        # (B,N,3) view of a (B,3,N) buffer: what models pass (xyz.permute(0,2,1))
    s_p = torch.sum(xp ** 2, -1).cpu().numpy().reshape(-1)
    out[f"B{B}_N{N}"]["permuted_view_vs_contiguous"] = int((s_p.view(np.int32) != s_t.view(np.int32)).sum())
    out[f"B{B}_N{N}"]["permuted_view_vs_(x2+y2)+z2"] = int((s_p.view(np.int32) != cands["(x2+y2)+z2"].view(np.int32)).sum())
    np.savez(os.path.join(ROOT, "gpurun_out", f"sumsq_B{B}_N{N}.npz"), gpu=s_t, perm=s_p)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_sumsq.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

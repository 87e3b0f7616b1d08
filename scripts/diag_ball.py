#!/usr/bin/env python
"""Diagnose ball-query mismatches against the reference's torch path ON THE GPU (cuBLAS square_distance).
Dumps, for every mismatching query, the points whose membership differs with the reference's sqrdist value,
r^2, and the value of the expanded formula this library assumes (emulated in float64->float32 fma steps)."""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import ref_models  # noqa: E402
from toothgroupnetwork_b200 import clouds  # noqa: E402
from toothgroupnetwork_b200 import pointnet2_utils as pn2  # noqa: E402


def f32(x):
    return np.float32(x)


def fma32(a, b, c):
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def formula(q, p):
    """d = -2*fma(qz,pz,fma(qy,py,qx*px)); d += |q|^2; d += |p|^2 with |v|^2 = (x*x+y*y)+z*z  (DESIGN.md 2)."""
    dot = fma32(q[2], p[2], fma32(q[1], p[1], f32(q[0] * p[0])))
    qq = f32(f32(f32(q[0] * q[0]) + f32(q[1] * q[1])) + f32(q[2] * q[2]))
    pp = f32(f32(f32(p[0] * p[0]) + f32(p[1] * p[1])) + f32(p[2] * p[2]))
    return f32(f32(f32(-2.0) * dot + qq) + pp)


def main():
    w = ref_models.World("reference")
    out = {}
    for B in (1, 2, 16):
        feats = torch.stack([clouds.arch_features(24000, s)[0] for s in range(B)]).cuda()
        xyz_t = feats[:, :3].permute(0, 2, 1).contiguous()
        new_xyz = pn2._take_rows(xyz_t.view(-1, 3), pn2._fps_batched(xyz_t, 1024)).view(B, 1024, 3)
        for r, K in ((0.025, 32), (0.05, 64)):
            with w:
                refpn = w.mod("external_libs.pointnet2_utils.pointnet2_utils")
                want = refpn.query_ball_point(r, K, xyz_t, new_xyz)
                sq = refpn.square_distance(new_xyz, xyz_t)
            got = pn2._ball_query(r, K, xyz_t, new_xyz, True)
            bad = (want != got).any(-1).nonzero()
            rec = {"mismatching_queries": int(bad.shape[0]), "detail": []}
            r2 = np.float32(r ** 2)
            # membership masks of the reference on GPU vs the emulated formula, whole matrix
            sq_cpu = sq.cpu().numpy()
            q_cpu, p_cpu = new_xyz.cpu().numpy(), xyz_t.cpu().numpy()
            for (b, s) in bad[:6].tolist():
                wi, gi = set(want[b, s].tolist()), set(got[b, s].tolist())
                for j in sorted(wi ^ gi)[:4]:
                    if j >= 24000:
                        continue
                    rec["detail"].append({"b": b, "s": s, "j": j, "in_ref": j in wi, "in_ours": j in gi,
                                          "ref_sqrdist": float(sq_cpu[b, s, j]), "r2": float(r2),
                                          "emulated_formula": float(formula(q_cpu[b, s], p_cpu[b, j])),
                                          "ref_bits": int(sq_cpu[b, s, j].view(np.int32)),
                                          "emu_bits": int(formula(q_cpu[b, s], p_cpu[b, j]).view(np.int32))})
            # how often does the emulated formula differ from cuBLAS at all (first cloud, first 64 queries)?
            diff = 0
            tot = 0
            for s in range(0, 1024, 64):
                for j in range(0, 24000, 37):
                    tot += 1
                    diff += int(formula(q_cpu[0, s], p_cpu[0, j]).view(np.int32) != sq_cpu[0, s, j].view(np.int32))
            rec["sampled_pairs"] = tot
            rec["sampled_pairs_formula_differs"] = diff
            out[f"B{B}_r{r}_K{K}"] = rec
    # what does torch.sum(x**2,-1) give against (x*x+y*y)+z*z ?
    x = xyz_t[0]
    s_t = torch.sum(x ** 2, -1).cpu().numpy()
    xc = x.cpu().numpy()
    s_e = ((xc[:, 0] * xc[:, 0] + xc[:, 1] * xc[:, 1]) + xc[:, 2] * xc[:, 2]).astype(np.float32)
    out["sumsq_differs"] = int((s_t.view(np.int32) != s_e.view(np.int32)).sum())
    mm = torch.matmul(new_xyz[0], xyz_t[0].t()).cpu().numpy()
    q, p = new_xyz[0].cpu().numpy().astype(np.float64), xyz_t[0].cpu().numpy().astype(np.float64)
    e1 = np.float32(np.float32(np.float32(q[:, None, 0] * p[None, :, 0]).astype(np.float64) + q[:, None, 1] * p[None, :, 1]).astype(np.float64)
                    + q[:, None, 2] * p[None, :, 2])
    out["matmul_differs_from_fma_chain"] = int((mm.view(np.int32) != e1.view(np.int32)).sum())
    e2 = np.float32(np.float32(np.float32(q[:, None, 2] * p[None, :, 2]).astype(np.float64) + q[:, None, 1] * p[None, :, 1]).astype(np.float64)
                    + q[:, None, 0] * p[None, :, 0])
    out["matmul_differs_from_reverse_fma_chain"] = int((mm.view(np.int32) != e2.view(np.int32)).sum())
    mmB = torch.matmul(new_xyz, xyz_t.permute(0, 2, 1))[0].cpu().numpy()
    out["batched_matmul_differs_from_single"] = int((mmB.view(np.int32) != mm.view(np.int32)).sum())
    out["batched_matmul_differs_from_fma_chain"] = int((mmB.view(np.int32) != e1.view(np.int32)).sum())
    out["batched_matmul_differs_from_reverse_fma_chain"] = int((mmB.view(np.int32) != e2.view(np.int32)).sum())
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_ball.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:6000])


if __name__ == "__main__":
    main()

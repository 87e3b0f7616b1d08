"""Seeded synthetic point clouds: the inputs every parity test, golden
fixture and bench line is quoted on.  CPU generators (torch.Generator) so that the same
seed gives the same cloud in this container and on the GPU box.

* ``cube``  -- xyz ~ U[-1,1]^3 ("synthetic random cloud", BASELINE config C1).
* ``dental_arch`` -- half-torus tube (centre line radius 0.6, tube radius 0.18, sigma 0.002
  noise) with analytic normals: surface-like density, so balls of r <= 0.1 hold >= K points the
  way real scans normalised per ``preprocess_data.py:48-50`` do (configs C2-C5).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def cloud_seed(rank: int, j: int) -> int:
    """Cloud j on rank r uses seed 1000*r + j."""
    return 1000 * int(rank) + int(j)


def cube(n: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(int(seed))
    return (torch.rand(n, 3, generator=g, dtype=torch.float32) * 2.0 - 1.0).contiguous()


def dental_arch(n: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (xyz (n,3), normal (n,3), label (n,) int64).  Labels: 16 equal u-bins ("teeth")
    on the upper half of the tube (sin v > 0), -1 ("gingiva") elsewhere."""
    g = torch.Generator().manual_seed(int(seed))
    u = torch.rand(n, generator=g, dtype=torch.float32) * math.pi
    v = torch.rand(n, generator=g, dtype=torch.float32) * (2.0 * math.pi)
    noise = torch.randn(n, 3, generator=g, dtype=torch.float32) * 0.002
    cu, su, cv, sv = torch.cos(u), torch.sin(u), torch.cos(v), torch.sin(v)
    centre = torch.stack([0.6 * cu, 0.6 * su - 0.3, torch.zeros_like(u)], 1)
    e_r = torch.stack([cu, su, torch.zeros_like(u)], 1)
    e_z = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    normal = cv[:, None] * e_r + sv[:, None] * e_z
    xyz = (centre + 0.18 * normal + noise).contiguous()
    label = torch.where(sv > 0, torch.clamp((u / math.pi * 16).long(), max=15), torch.full_like(u, -1).long())
    return xyz, normal.contiguous(), label


def arch_features(n: int, seed: int = 0) -> torch.Tensor:
    """(1, 6, n) channel-first feature tensor [xyz; normal] as the models consume it
    (``models/modules/grouping_network_module.py:16-23``)."""
    xyz, normal, _ = dental_arch(n, seed)
    return torch.cat([xyz, normal], 1).t().contiguous().unsqueeze(0)


def with_duplicates(xyz: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """Every vertex twice, shuffled: the FPS / kNN tie-break stress input."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    both = torch.cat([xyz, xyz], 0)
    return both[torch.randperm(both.shape[0], generator=g)].contiguous()

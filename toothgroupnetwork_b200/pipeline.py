"""Host-to-host pipelining of a set-abstraction module over a large batch of clouds.

The operators launch on the *current* CUDA stream (and the FPS workspace is stream-ordered), so a
batch that lives in pinned host memory can be processed in chunks on a few streams: the H2D copy
of chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i (PCIe is full duplex),
and kernels of different chunks overlap each other's latency-bound phases on the SMs.

    pipe = HostPipeline(sa_module, chunk_clouds=148, n_streams=4)
    pipe(host_feats, out_xyz_host, out_points_host)      # all pinned; returns after a full sync

``host_feats`` is the reference's model input layout ``(B, C, N)`` with xyz in channels 0..2
(``models/modules/pointnet_pp.py:43-47``: ``l0_xyz = xyz[:, :3, :]``).
"""
from __future__ import annotations

from typing import List

import torch


class HostPipeline:
    def __init__(self, module: torch.nn.Module, chunk_clouds: int = 148, n_streams: int = 4):
        self.module = module
        self.chunk = int(chunk_clouds)
        self.streams: List[torch.cuda.Stream] = [torch.cuda.Stream() for _ in range(max(1, int(n_streams)))]

    @torch.no_grad()
    def __call__(self, host_feats: torch.Tensor, out_xyz_host: torch.Tensor, out_points_host: torch.Tensor) -> None:
        B = host_feats.shape[0]
        main = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(main)
        for k, lo in enumerate(range(0, B, self.chunk)):
            hi = min(B, lo + self.chunk)
            s = self.streams[k % len(self.streams)]
            with torch.cuda.stream(s):
                d = host_feats[lo:hi].to("cuda", non_blocking=True)
                new_xyz, new_points = self.module(d[:, :3].contiguous(), d)
                out_xyz_host[lo:hi].copy_(new_xyz, non_blocking=True)
                out_points_host[lo:hi].copy_(new_points, non_blocking=True)
        for s in self.streams:
            main.wait_stream(s)

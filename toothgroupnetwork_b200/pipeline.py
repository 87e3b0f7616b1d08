"""Host-to-host pipelining of a set-abstraction module over a large batch of clouds.

The operators launch on the *current* CUDA stream (and the FPS workspace is stream-ordered), so a
batch that lives in pinned host memory can be processed in chunks: the H2D copies of all chunks are
queued back to back on a copy stream, each chunk computes on one of a few compute streams as soon
as it has landed, and results return on a third stream (PCIe is full duplex).

    pipe = HostPipeline(sa_module, chunk_clouds=148, n_streams=2)
    pipe(host_feats, out_xyz_host, out_points_host)      # all pinned; returns after a full sync

``host_feats`` is the reference's model input layout ``(B, C, N)`` with xyz in channels 0..2
(``models/modules/pointnet_pp.py:43-47``: ``l0_xyz = xyz[:, :3, :]``).
"""
from __future__ import annotations

from typing import List

import torch

from . import pointnet2_utils as pn2


class HostPipeline:
    """``n_streams`` compute streams; the H2D and D2H copies run on two dedicated streams so that
    the host-to-device engine streams the whole batch back to back (it is the bottleneck: the
    kernels of a chunk take less time than its PCIe transfer) while chunks compute as they land."""

    def __init__(self, module: torch.nn.Module, chunk_clouds: int = 148, n_streams: int = 2):
        self.module = module
        self.chunk = int(chunk_clouds)
        self.streams: List[torch.cuda.Stream] = [torch.cuda.Stream() for _ in range(max(1, int(n_streams)))]
        self.copy_in = torch.cuda.Stream()
        self.copy_out = torch.cuda.Stream()

    @torch.no_grad()
    def __call__(self, host_feats: torch.Tensor, out_xyz_host: torch.Tensor, out_points_host: torch.Tensor) -> None:
        B = host_feats.shape[0]
        main = torch.cuda.current_stream()
        self.copy_in.wait_stream(main)
        self.copy_out.wait_stream(main)
        for s in self.streams:
            s.wait_stream(main)
        spans = [(lo, min(B, lo + self.chunk)) for lo in range(0, B, self.chunk)]
        # FPS shape for the clouds resident on the GPU (all compute streams), not for one chunk
        saved_mode = pn2._fps_mode
        pn2.set_fps_mode(pn2.fps_mode_for_clouds_in_flight(min(B, self.chunk * len(self.streams)), host_feats.shape[2]))
        landed = []
        with torch.cuda.stream(self.copy_in):                  # every H2D copy queued up front, back to back
            for lo, hi in spans:
                d = host_feats[lo:hi].to("cuda", non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_in)
                landed.append((d, ev))
        for k, (lo, hi) in enumerate(spans):
            d, ev = landed[k]
            s = self.streams[k % len(self.streams)]
            s.wait_event(ev)
            d.record_stream(s)
            with torch.cuda.stream(s):
                new_xyz, new_points = self.module(d[:, :3].contiguous(), d)
                done = torch.cuda.Event()
                done.record(s)
            self.copy_out.wait_event(done)
            new_xyz.record_stream(self.copy_out)
            new_points.record_stream(self.copy_out)
            with torch.cuda.stream(self.copy_out):
                out_xyz_host[lo:hi].copy_(new_xyz, non_blocking=True)
                out_points_host[lo:hi].copy_(new_points, non_blocking=True)
        landed.clear()
        pn2.set_fps_mode(saved_mode)
        main.wait_stream(self.copy_out)
        for s in self.streams:
            main.wait_stream(s)

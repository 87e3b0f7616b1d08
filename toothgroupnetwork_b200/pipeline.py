"""Host-to-host pipelining of a set-abstraction module over a large batch of clouds.

The operators launch on the *current* CUDA stream (and the FPS workspace is stream-ordered), so a
batch that lives in pinned host memory can be processed in chunks: the H2D copies of all chunks are
queued back to back on a copy stream, groups of chunks compute on one of a few compute streams as
soon as they have landed, and results return on a third stream (PCIe is full duplex).

    pipe = HostPipeline(sa_module, chunk_clouds=148, n_streams=2)
    pipe(host_feats, out_xyz_host, out_points_host)      # all pinned; returns after a full sync

``host_feats`` is the reference's model input layout ``(B, C, N)`` with xyz in channels 0..2
(``models/modules/pointnet_pp.py:43-47``: ``l0_xyz = xyz[:, :3, :]``).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import pointnet2_utils as pn2


class HostPipeline:
    """H2D copies run back to back on a dedicated stream in chunks of ``chunk_clouds`` into ONE device
    buffer; compute runs on ``n_streams`` streams over GROUPS of consecutive chunks (``groups``: chunks
    per group in order, the last but one repeating, the last entry being the size of the final group) as soon as the last chunk of a group has landed;
    results return on a third stream.  Measured on B200 (1184 clouds of 24k points, scripts/gpu_e2e_sweep.sh):
    one chunk per group is best (5.9e7 sampled points/s); larger groups start later than they gain in
    kernel efficiency ((2,3,2,1): 5.5e7, (2,5,1): 4.6e7), so that is the default."""

    def __init__(self, module: torch.nn.Module, chunk_clouds: int = 148, n_streams: int = 2,
                 groups: Sequence[int] = (1,), fps_mode: int = None):
        self.module = module
        self.chunk = int(chunk_clouds)
        self.groups = tuple(int(g) for g in groups) or (1,)
        self.fps_mode = fps_mode          # None: shape by the clouds in flight; else the tgn_furthestsampling mode of every chunk
        self.streams: List[torch.cuda.Stream] = [torch.cuda.Stream() for _ in range(max(1, int(n_streams)))]
        self.copy_in = torch.cuda.Stream()
        self.copy_out = torch.cuda.Stream()
        self._dev = None

    def _plan(self, B: int):
        """[(lo, hi)] of the chunks and [(first_chunk, last_chunk)] of the compute groups."""
        spans = [(lo, min(B, lo + self.chunk)) for lo in range(0, B, self.chunk)]
        body, tail = self.groups[:-1], self.groups[-1]
        plan, k, i = [], 0, 0
        while k < len(spans):
            remaining = len(spans) - k
            if remaining <= tail or not body:
                n = min(remaining, tail)
            else:                                   # body entries in order, the last one repeating; keep the tail group
                n = max(1, min(body[min(i, len(body) - 1)], remaining - tail))
                i += 1
            plan.append((k, k + n - 1))
            k += n
        return spans, plan

    @torch.no_grad()
    def __call__(self, host_feats: torch.Tensor, out_xyz_host: torch.Tensor, out_points_host: torch.Tensor) -> None:
        B = host_feats.shape[0]
        main = torch.cuda.current_stream()
        if self._dev is None or self._dev.shape != host_feats.shape:
            self._dev = torch.empty(host_feats.shape, dtype=host_feats.dtype, device="cuda")
        dev = self._dev
        self.copy_in.wait_stream(main)
        self.copy_out.wait_stream(main)
        for s in self.streams:
            s.wait_stream(main)
        spans, plan = self._plan(B)
        landed = []
        with torch.cuda.stream(self.copy_in):                  # every H2D copy queued up front, back to back
            for lo, hi in spans:
                dev[lo:hi].copy_(host_feats[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_in)
                landed.append(ev)
        saved_mode = pn2._fps_mode
        for gi, (k0, k1) in enumerate(plan):
            lo, hi = spans[k0][0], spans[k1][1]
            s = self.streams[gi % len(self.streams)]
            s.wait_event(landed[k1])
            # FPS shape for the clouds resident on the GPU (this group and its neighbour on the other stream)
            pn2.set_fps_mode(self.fps_mode if self.fps_mode is not None else
                             pn2.fps_mode_for_clouds_in_flight(min(B, (hi - lo) * len(self.streams)), host_feats.shape[2]))
            with torch.cuda.stream(s):
                d = dev[lo:hi]
                new_xyz, new_points = self.module(d[:, :3].contiguous(), d)
                if not new_xyz.is_contiguous():        # the modules return the reference's permuted view of (B,S,3)
                    new_xyz = pn2.transpose_last2(new_xyz.permute(0, 2, 1))
                done = torch.cuda.Event()
                done.record(s)
            self.copy_out.wait_event(done)
            new_xyz.record_stream(self.copy_out)
            new_points.record_stream(self.copy_out)
            with torch.cuda.stream(self.copy_out):
                out_xyz_host[lo:hi].copy_(new_xyz, non_blocking=True)
                out_points_host[lo:hi].copy_(new_points, non_blocking=True)
        pn2.set_fps_mode(saved_mode)
        main.wait_stream(self.copy_out)
        main.wait_stream(self.copy_in)
        for s in self.streams:
            main.wait_stream(s)

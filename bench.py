#!/usr/bin/env python
"""bench.py -- sampled-points/sec of FPS + ball query + grouped shared-MLP on 24k-point clouds.

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2(i)"): PointNet++ set-abstraction level 1,
``PointNetSetAbstraction(npoint=1024, radius=0.1, nsample=32, in_channel=6+3, mlp=[32,32,64])``,
eval-mode BatchNorm, forward only, on a batch of synthetic dental-arch clouds of 24 000 points
(xyz + normals).  One step = one pass over one batch of ``--clouds`` clouds per GPU:
    FPS 24000->1024  ->  ball query (r=0.1, K=32)  ->  fused gather+MLP(9->32->32->64)+max.
metric = sampled points per second = clouds * 1024 / time, whole job (all ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clouds B]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

* ``value``: inputs resident in HBM, CUDA-event timed, max over ranks.
* ``e2e``: the same step through the public module API with HOST (pinned) inputs: H2D copy of the
  (B,6,N) feature tensor and D2H copy of both outputs inside the timed region.
* ``roofline``: the dominant kernel (FPS) -- algorithmic bytes 20*(M-1)*N per cloud over its
  CUDA-event time inside the timed region, against the measured HBM copy peak.
* ``cpu_baseline`` / ``--impl reference``: the oracle's port of the reference's CPU-capable
  formulation of the same path, timed on the host cores on a bounded sample of the workload.
Inputs are larger than L2 (B*24000*6*4 bytes = 682 MB at the default B = 1184); no explicit flush.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, NPOINT, RADIUS, NSAMPLE = 24000, 1024, 0.1, 32
MLP = [32, 32, 64]
METRIC = "sampled-points/sec (FPS+ballq+group-MLP, 24k-pt cloud)"
UNIT = "sampled points/s"
WORKLOAD = "pointnet++ SA1 forward: FPS 24000->1024, ball query r=0.1 K=32, group-MLP 9->[32,32,64], eval BN"
# dram__bytes_read.sum + dram__bytes_write.sum of fps_bucket_sort_kernel<1> + fps_bucket_kernel<128> in one
# `ncu --set full` capture of this bench at 1184 clouds (profiles/r1d_ncu_full_raw.csv): 18.12 GB per launch pair
NCU_FPS_DRAM_BYTES_PER_CLOUD = (16.000872e9 + 1.228595e9 + 0.353967e9 + 0.538954e9) / 1184


def host_cores() -> int:
    """Host threads this process may actually use (cgroup / affinity aware), not the machine total."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:   # cgroup v2 CPU quota
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_clouds(rank: int, count: int) -> torch.Tensor:
    """(count, 6, N) feature tensors [xyz; normal]; cloud j on rank r has seed 1000*r + j.
    To keep start-up short only 8 distinct clouds are synthesised per rank and tiled (FPS /
    ball query / MLP cost does not depend on which cloud is processed)."""
    from toothgroupnetwork_b200 import clouds
    base = [clouds.arch_features(N_POINTS, clouds.cloud_seed(rank, j))[0] for j in range(min(count, 8))]
    return torch.stack([base[j % len(base)] for j in range(count)]).contiguous()


def build_module(device):
    from toothgroupnetwork_b200 import pointnet2_utils as pn2
    torch.manual_seed(0)
    sa = pn2.PointNetSetAbstraction(NPOINT, RADIUS, NSAMPLE, 9, MLP, False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for bn in sa.mlp_bns:     # non-trivial eval statistics
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
    return sa.to(device).eval()


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(feats: torch.Tensor, sa_cpu_layers, threads: int):
    """The oracle's port of the reference's CPU-capable formulation on ``feats`` (b,6,N):
    torch-loop FPS with start 0 (pointnet2_utils.py:103-118), query_ball_point (:120-144), gather,
    conv1x1+BN+ReLU x3, max (:227-237).  Clouds are spread over ``threads`` host threads for the
    sampling / search, the MLP runs with torch's intra-op threads."""
    from oracle import oracle
    xyz = feats[:, :3].permute(0, 2, 1).contiguous().numpy()
    _, new_np, gi_np = oracle.sa_sample_and_search_batch(xyz, NPOINT, RADIUS, NSAMPLE)   # one cloud per OpenMP thread
    new_xyz = torch.from_numpy(new_np)
    gidx = torch.from_numpy(gi_np)
    pts = feats.permute(0, 2, 1)
    xyz_t = torch.from_numpy(xyz)
    grouped = torch.cat([oracle.index_points(xyz_t, gidx) - new_xyz.unsqueeze(2), oracle.index_points(pts, gidx)], -1)
    h = grouped.permute(0, 3, 2, 1)
    for p in sa_cpu_layers:
        h = oracle._conv_bn_relu(h, p, False)
    return new_xyz, h.max(dim=2)[0]


def cpu_layers_of(sa):
    from oracle import oracle
    out = []
    for c, b in zip(sa.mlp_convs, sa.mlp_bns):
        out.append(oracle.MlpParams(c.weight.detach().cpu().reshape(c.weight.shape[0], -1), c.bias.detach().cpu(),
                                    b.weight.detach().cpu(), b.bias.detach().cpu(), b.running_mean.detach().cpu(),
                                    b.running_var.detach().cpu(), b.eps))
    return out


def time_cpu(sa, sample_clouds: int, steps: int, warmup: int):
    cores = host_cores()
    torch.set_num_threads(cores)
    feats = make_clouds(0, sample_clouds)
    layers = cpu_layers_of(sa)
    for _ in range(warmup):
        cpu_reference_step(feats, layers, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(feats, layers, cores)
    dt = (time.perf_counter() - t0) / steps
    return sample_clouds * NPOINT / dt, dt, cores


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clouds", type=int, default=1184, help="clouds per GPU per step (8 per SM)")
    ap.add_argument("--cpu-clouds", type=int, default=0, help="clouds in the CPU sample (default: host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=148, help="clouds per pipelined chunk of the end-to-end measurement")
    ap.add_argument("--e2e-streams", type=int, default=2)
    ap.add_argument("--e2e-groups", default="1", help="chunks per compute group of the end-to-end pipeline")
    ap.add_argument("--fps-mode", type=int, default=0, help="0 auto; 100*G+CS resident shape, -2 bucket, -(10+W) bucket with W warps per cloud (experiments)")
    ap.add_argument("--ball-path", type=int, default=0, help="0 auto, 4 index-order tile scan, 8 uniform grid (experiments)")
    ap.add_argument("--sa-engine", type=int, default=0, help="0 auto, 1 fp32 CUDA cores, 2 tcgen05 (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from toothgroupnetwork_b200 import sharding
    rank, world, local = sharding.env_rank_world()
    config = {"workload": WORKLOAD, "clouds_per_gpu": args.clouds, "points_per_cloud": N_POINTS, "npoint": NPOINT,
              "radius": RADIUS, "nsample": NSAMPLE, "mlp": MLP, "parallelism": f"mesh-sharded x{world}",
              "l2": "inputs larger than L2, no flush"}

    if args.impl == "reference":
        # Reference arm: CPU, rank 0 only, bounded sample of the same workload.
        if rank != 0:
            return
        sa = build_module("cpu")
        cores = host_cores()
        sample = args.cpu_clouds or cores
        value, dt, cores = time_cpu(sa, sample, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{sample} clouds of {N_POINTS} points per step (one per host thread)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world, local = sharding.init("nccl")
    from toothgroupnetwork_b200 import _lib as L
    from toothgroupnetwork_b200 import pointnet2_utils as pn2

    sa = build_module(device)
    pn2.set_sa_engine(args.sa_engine)
    pn2.set_ball_path(args.ball_path)
    B = args.clouds
    host_feats = make_clouds(rank, B).pin_memory()
    feats = host_feats.to(device)
    xyz = feats[:, :3].contiguous()
    folded = sa._folded.update(sa.mlp_convs, sa.mlp_bns)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def staged_step(events=None):
        """Same kernels as sa.forward, with events between the three stages."""
        xyz_t = pn2.transpose_last2(xyz)
        feats_t = pn2.transpose_last2(feats)
        if events: events[0].record()
        fps = pn2._fps_batched(xyz_t, NPOINT, args.fps_mode)
        if events: events[1].record()
        new_xyz_t = pn2._take_rows(xyz_t.view(-1, 3), fps).view(B, NPOINT, 3)
        gidx = pn2._ball_query(RADIUS, NSAMPLE, xyz_t, new_xyz_t, False)
        if events: events[2].record()
        out = torch.empty((B, MLP[-1], NPOINT), dtype=torch.float32, device=device)
        pn2.sa_group_mlp_max(xyz_t, feats_t, new_xyz_t, gidx, True, folded, out, 0)
        if events: events[3].record()
        return pn2.transpose_last2(new_xyz_t), out

    # -------- parity smoke on this rank: kernel outputs vs module API (bitwise) ----------------
    with torch.no_grad():
        a = staged_step()
        b = sa(xyz, feats)
    parity_ok = bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))

    # -------- device-resident timing ----------------------------------------------------------------
    with torch.no_grad():
        for _ in range(args.warmup):
            staged_step()
        stage_events = [[ev() for _ in range(4)] for _ in range(args.steps)]
        t_start, t_end = ev(), ev()
        sharding.barrier()
        torch.cuda.synchronize()
        launches0 = L.launch_count()
        with ClockSampler(local) as clk:
            t_start.record()
            for s in range(args.steps):
                staged_step(stage_events[s])
            t_end.record()
            torch.cuda.synchronize()
        launches = L.launch_count() - launches0
        sharding.barrier()
    local_s = t_start.elapsed_time(t_end) / 1e3
    total_s = sharding.max_over_ranks(local_s, device)
    fps_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in stage_events]))
    ball_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in stage_events]))
    mlp_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in stage_events]))
    value = world * B * NPOINT * args.steps / total_s

    # -------- end to end through the public API with host buffers -----------------------------------
    out_xyz_host = torch.empty((B, 3, NPOINT), dtype=torch.float32).pin_memory()
    out_pts_host = torch.empty((B, MLP[-1], NPOINT), dtype=torch.float32).pin_memory()

    from toothgroupnetwork_b200.pipeline import HostPipeline
    pipe = HostPipeline(sa, chunk_clouds=args.e2e_chunk, n_streams=args.e2e_streams,
                        groups=[int(g) for g in args.e2e_groups.split(",")])

    def e2e_step():
        # public API on host buffers: chunks of the batch go H2D -> module.forward -> D2H on a few
        # streams, so copies overlap kernels (every byte still crosses PCIe inside the timed region)
        pipe(host_feats, out_xyz_host, out_pts_host)

    with torch.no_grad():
        for _ in range(args.warmup):
            e2e_step()
        e0, e1 = ev(), ev()
        sharding.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
    e2e_s = sharding.max_over_ranks(e0.elapsed_time(e1) / 1e3, device)
    e2e_value = world * B * NPOINT * args.steps / e2e_s

    # -------- the one collective of the run: per-rank metric records ----------------------------------
    records = sharding.gather_metrics({"sampled_points": B * NPOINT * args.steps, "clouds": B * args.steps, "seconds": local_s,
                                       "parity_ok": float(parity_ok), "launches": launches}, device)
    agg = sharding.reduce_metrics(records)

    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    fps_alg_bytes = 20.0 * (NPOINT - 1) * N_POINTS * B            # per FPS launch
    achieved = fps_alg_bytes / (fps_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "stage_ms": {"fps": fps_ms, "ball_query": ball_ms, "group_mlp": mlp_ms},
        "roofline": {"kernel": "fps_bucket_kernel (+ its sort prologue; one event pair brackets both)", "bound": "hbm",
                     "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": NCU_FPS_DRAM_BYTES_PER_CLOUD * B, "peak_kind": peak_kind,
                     "dram_achieved_gbs": NCU_FPS_DRAM_BYTES_PER_CLOUD * B / (fps_ms * 1e-3) / 1e9,
                     "dram_frac": NCU_FPS_DRAM_BYTES_PER_CLOUD * B / (fps_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "note": "achieved = algorithmic 20*(M-1)*N*B bytes / FPS time; frac > 1 because the exact "
                             "bucket pruning never touches ~97% of those point-updates (profiles/r1_summary.md); dram_frac is the "
                             "share of the measured HBM peak the kernel's real (random, 1.3 KB-granular) traffic reaches. "
                             "traffic = ncu dram read+write of the same launches (profiles/r1d_ncu_full_raw.csv), scaled per cloud"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host_feats.numel() * 4),
                "d2h_bytes_per_step": int((out_xyz_host.numel() + out_pts_host.numel()) * 4)},
        "gpu_launches": int(agg["launches"]),
        "clocks": clk.summary(),
        "parity_ok": bool(agg["parity_ok"]),
    }
    if not args.no_cpu_baseline:
        cores = host_cores()
        sample = args.cpu_clouds or cores
        cv, cdt, cores = time_cpu(sa.cpu(), sample, 2, 1)
        line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} clouds of {N_POINTS} points per step (one per host thread), 2 timed steps"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()

"""Mesh sharding over ranks.

Every scan / crop is an independent unit -- all operators are segment-local -- so the data path
shards by mesh with no collective: rank r owns meshes {j : j mod world == r}.  The only exchange
is the reduction of per-rank metric records at the end of a run (one all_gather over NCCL on the
GPU box, gloo in the CPU tests), plus the barrier / max-over-ranks used for timing.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist

METRIC_KEYS = ("sampled_points", "clouds", "seconds", "parity_ok", "launches", "numa_node", "e2e_seconds")


def env_rank_world() -> tuple:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str = "nccl") -> tuple:
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def owned(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the meshes rank ``rank`` processes."""
    return list(range(rank, n_items, world))


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device) -> float:
    """Timing rule: a multi-GPU duration is the max over ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_metrics(local: Dict[str, float], device: torch.device) -> List[Dict[str, float]]:
    """all_gather of one fixed-size record per rank (the single collective of a run)."""
    rec = torch.tensor([float(local.get(k, 0.0)) for k in METRIC_KEYS], dtype=torch.float64, device=device)
    if not dist.is_initialized():
        return [dict(zip(METRIC_KEYS, rec.tolist()))]
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    return [dict(zip(METRIC_KEYS, o.tolist())) for o in out]


def reduce_metrics(records: Sequence[Dict[str, float]]) -> Dict[str, float]:
    """Whole-job aggregate: points and clouds add up, time is the slowest rank, parity must hold
    everywhere."""
    return {
        "sampled_points": sum(r["sampled_points"] for r in records),
        "clouds": sum(r["clouds"] for r in records),
        "seconds": max(r["seconds"] for r in records),
        "parity_ok": float(all(r["parity_ok"] >= 1.0 for r in records)),
        "launches": sum(r["launches"] for r in records),
    }

"""Replica-parallel plumbing for the multi-GPU runs (one process per GPU, torchrun).

The SVDQuant linear path has no exchange step: images are independent (SURVEY.md section 8e, F9;
the reference's own multi-GPU story is N processes striding the prompt list, app/flux.1/t2i/
evaluate.py:30-39,68-69).  So the only collectives are control-plane ones, outside the timed kernel
path: broadcast of the job description from rank 0, gather of per-rank results, and a MAX
reduction of the per-rank device time so that throughput = total units / slowest rank.
Works with backend "nccl" (GPUs, NVLink) and "gloo" (CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None, device: torch.device | None = None) -> None:
    rank, ws, _ = world()
    if ws <= 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)


def shutdown() -> None:
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def stripe(units, rank: int, world_size: int):
    """Units (images / prompts / seeds) owned by ``rank``: i % world == rank, the reference's
    --chunk-start/--chunk-step striding."""
    return [u for i, u in enumerate(units) if i % world_size == rank]


def broadcast_job(job, src: int = 0):
    """Rank ``src``'s picklable job description to everyone."""
    if not dist.is_initialized():
        return job
    box = [job if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def gather_results(result, dst: int = 0):
    """List of every rank's picklable result on ``dst`` (None elsewhere)."""
    if not dist.is_initialized():
        return [result]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(result, out, dst=dst)
    return out


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    """Slowest rank's time: multi-GPU numbers are timed on the device and reduced with MAX."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()

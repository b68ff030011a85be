"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch; gloo in the CPU
tests).  The reference has no multi-GPU code — this is new design, two modes:

* "replicated": tables replicated, the POINT BATCH is sharded (`shard_range`); every per-point gradient already
  carries 1/N_global, so one sum all-reduce of the flat gradient buffer (tables + 1 377 decoder floats) yields
  exactly the single-GPU gradient of the global batch.
* "spatial" (BASELINE config 5): every rank owns a spatial octree block and the samples that fall inside it; table
  rows are private to their owner, only the decoder segment is all-reduced.  (Rows on shared block faces would
  need a boundary exchange; blocks used here are disjoint — see DESIGN.md.)
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  -> (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of an n-point batch for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t


def max_over_ranks(value: float, device) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def barrier(device=None):
    if dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index or 0])
        else:
            dist.barrier()

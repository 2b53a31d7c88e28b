"""Data parallelism for the hot path: one process per GPU, trajectories sharded,
model replicated.  Rollout needs no collective; training has exactly one exchange per step — the
all-reduce of the flat gradient buffer the backward kernels accumulate into (NCCL over NVLink)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n independent trajectories for `rank`."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_trajectories(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e]


def allreduce_mean_(flat: torch.Tensor, world: int | None = None) -> torch.Tensor:
    """In-place mean over ranks of a flat gradient buffer (sum, then x 1/world — MSE 'mean'
    semantics are preserved when local batches are equal)."""
    if not dist.is_initialized():
        return flat
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    return flat


def train_step(model, optimizer, loss_fn, data, target, ib):
    """The reference inner loop (train/train_temporal.py:254-258) + the DP gradient exchange."""
    optimizer.zero_grad(set_to_none=True)
    out = model(data, ib)
    loss = loss_fn(out, target)
    loss.backward()
    eng = getattr(model, "_sea_engine", None) or model.engine()
    allreduce_mean_(eng.flat_grad())
    optimizer.step()
    return loss

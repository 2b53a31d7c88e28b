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
    """In-place mean over ranks of a flat gradient buffer (MSE 'mean' semantics are preserved when local
    batches are equal).  NCCL averages inside the collective (no extra pass over the buffer); other
    backends sum, then scale."""
    if not dist.is_initialized():
        return flat
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return flat
    if flat.is_cuda and dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    return flat


# Overlap pays once the overlapped bucket is worth a second collective launch: on 2 B200s (scripts/dp_train_bench.py,
# profiles/r1d_dp.md) both configs gain (cylinder_flow: 128 MiB bucket, 2.61 -> 2.45 ms; multiphase_flow: 512 MiB,
# 4.31 -> 3.69 ms); toy models keep the single collective.
OVERLAP_MIN_BYTES = 64 << 20

_overlap_state = {}


def _overlap_handles(dev: torch.device):
    """(side stream, cudaEvent_t) per device for the overlapped exchange."""
    import ctypes as C

    from ._lib import check, lib
    key = (dev.type, dev.index)
    if key not in _overlap_state:
        ev = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.sea_event_create(C.byref(ev)), "event_create")
            _overlap_state[key] = (torch.cuda.Stream(device=dev), ev)
    return _overlap_state[key]


def exchange_gradients(eng, armed_event=None, world: int | None = None) -> None:
    """Mean over ranks of the engine's flat gradient buffer.  With `armed_event` (recorded by the backward once
    the stream-MLP weight gradients are final) the tail bucket is reduced on a side stream, concurrently with
    the rest of the backward; the head bucket follows on the compute stream."""
    if not dist.is_initialized():
        return
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return
    flat = eng.flat_grad()
    k = eng.mlp_grad_offset()
    if armed_event is None or not flat.is_cuda or k >= flat.numel():
        allreduce_mean_(flat, world)
        return
    import ctypes as C

    from ._lib import check, lib
    comm, ev = _overlap_handles(flat.device)
    check(lib.sea_stream_wait_event(C.c_void_p(comm.cuda_stream), ev), "stream_wait_event")
    with torch.cuda.stream(comm):
        allreduce_mean_(flat[k:], world)
    if k > 0:
        allreduce_mean_(flat[:k], world)
    torch.cuda.current_stream(flat.device).wait_stream(comm)


def train_step(model, optimizer, loss_fn, data, target, ib, overlap: bool | None = None):
    """The reference inner loop (train/train_temporal.py:254-258) + the DP gradient exchange.  The exchange of
    the stream-MLP bucket overlaps the tail of the backward when `overlap` is True, or when it
    is None (default) and that bucket holds at least OVERLAP_MIN_BYTES."""
    eng = getattr(model, "_sea_engine", None) or model.engine()
    optimizer.zero_grad(set_to_none=True)
    out = model(data, ib)
    loss = loss_fn(out, target)
    armed = None
    if overlap is None:
        overlap = out.is_cuda and (eng.flat_grad().numel() - eng.mlp_grad_offset()) * 4 >= OVERLAP_MIN_BYTES
    if overlap and dist.is_initialized() and dist.get_world_size() > 1 and out.is_cuda:
        from ._lib import lib
        _, armed = _overlap_handles(out.device)
        lib.sea_temporal_backward_milestone(armed)
    loss.backward()
    if armed is not None:
        from ._lib import lib
        if lib.sea_temporal_backward_milestone_pending():   # the backward did not run on this thread / path
            lib.sea_temporal_backward_milestone(None)
            armed = None
    exchange_gradients(eng, armed)
    optimizer.step()
    return loss

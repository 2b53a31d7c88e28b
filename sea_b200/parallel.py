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


# ---- overlapped, bucketed gradient exchange --------------------------------------------------------------------
# The flat gradient buffer is ordered by the moment the backward FINISHES each group of weight gradients
# (TemporalEngine.grad_buckets(), sea_temporal_desc.bwd_events): final norms | stream MLP + proj | ln.exp.2 + TIPI |
# exchange | self-attention, then the small reduction-produced gradients.  Every group's event lets its bucket go out
# on a side stream while the rest of the backward still runs; groups below MIN_BUCKET_BYTES ride with their successor
# (one more collective costs ~20 us of launch + protocol latency).  With grad_dtype="bf16" the weight-gradient GEMMs
# mirror their result as bf16 (second store of the same epilogue) and the exchange moves half the bytes; the fused
# AdamW then reads the averaged bf16 bucket directly.
MIN_BUCKET_BYTES = 24 << 20
OVERLAP_MIN_BYTES = 64 << 20    # models whose whole gradient is smaller keep the single collective

_overlap_state = {}


def _overlap_handles(dev: torch.device, n_events: int):
    """(side stream, [cudaEvent_t] * n_events) per device for the overlapped exchange."""
    import ctypes as C

    from ._lib import check, lib
    key = (dev.type, dev.index)
    if key not in _overlap_state or len(_overlap_state[key][1]) < n_events:
        with torch.cuda.device(dev):
            evs = []
            for _ in range(n_events):
                ev = C.c_void_p()
                check(lib.sea_event_create(C.byref(ev)), "event_create")
                evs.append(ev)
            _overlap_state[key] = (torch.cuda.Stream(device=dev, priority=-1), evs)   # high priority: the collective's CTAs
            # are scheduled as soon as SMs free up under the persistent backward GEMMs
    return _overlap_state[key]


def plan_buckets(groups, elem_bytes: int, min_bytes: int = MIN_BUCKET_BYTES):
    """Merge the per-event weight groups [(begin, end)] (exchange order) into collectives of at least `min_bytes`:
    returns [(begin, end, k)] where k is the LAST group inside the bucket — the event the collective has to wait for."""
    out, begin = [], None
    for k, (b, e) in enumerate(groups):
        if e <= b:
            continue
        if begin is None:
            begin = b
        if (e - begin) * elem_bytes >= min_bytes:
            out.append((begin, e, k))
            begin = None
        else:
            last = (begin, e, k)
    if begin is not None:                      # leftover tail: joins the previous bucket (which then waits for its event)
        if out:
            pb, _, _ = out.pop()
            out.append((pb, last[1], last[2]))
        else:
            out.append(last)
    return out


def exchange_gradients(eng, events=None, world: int | None = None, grad_dtype: str = "f32") -> bool:
    """Mean over ranks of the engine's gradients.  Without `events` (CPU / gloo, toy models): one collective over the
    flat fp32 buffer.  With `events` (the handles given to the backward as desc.bwd_events): one collective per bucket
    on a side stream, each released by its event; grad_dtype="bf16" exchanges the bf16 twin of the weight gradients
    (the small gradients always travel in fp32).  Returns True when the averaged WEIGHT gradients live in the bf16
    twin (eng.flat_grad_bf16()) rather than in flat_grad() / param.grad."""
    if not dist.is_initialized():
        return False
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return False
    flat = eng.flat_grad()
    if events is None or not flat.is_cuda:
        allreduce_mean_(flat, world)
        return False
    import ctypes as C

    from ._lib import check, lib
    use_bf16 = grad_dtype == "bf16"
    buckets = eng.grad_buckets()
    weights, small = buckets[:-1], buckets[-1]
    buf = eng.flat_grad_bf16() if use_bf16 else flat
    comm, _ = _overlap_handles(flat.device, len(events))
    cs = C.c_void_p(comm.cuda_stream)
    plan = plan_buckets(weights, 2 if use_bf16 else 4)
    with torch.cuda.stream(comm):
        for b, e, k in plan:
            check(lib.sea_stream_wait_event(cs, events[k]), "stream_wait_event")
            allreduce_mean_(buf[b:e], world)
        if small[1] > small[0]:
            check(lib.sea_stream_wait_event(cs, events[-1]), "stream_wait_event")
            allreduce_mean_(flat[small[0]:small[1]], world)
    torch.cuda.current_stream(flat.device).wait_stream(comm)
    return use_bf16


def train_step(model, optimizer, loss_fn, data, target, ib, overlap: bool | None = None, grad_dtype: str | None = None):
    """The reference inner loop (train/train_temporal.py:254-258) + the DP gradient exchange, bucket by bucket behind
    the backward when `overlap` is True, or when it is None (default) and the gradient is at least
    OVERLAP_MIN_BYTES.  grad_dtype: "bf16" (default for a bf16 engine) or "f32" buckets."""
    eng = getattr(model, "_sea_engine", None) or model.engine()
    optimizer.zero_grad(set_to_none=True)
    out = model(data, ib)
    loss = loss_fn(out, target)
    distributed = dist.is_initialized() and dist.get_world_size() > 1 and out.is_cuda
    if grad_dtype is None:
        grad_dtype = "bf16" if getattr(eng, "precision", "bf16") == "bf16" else "f32"
    if overlap is None:
        overlap = out.is_cuda and eng.flat_grad().numel() * 4 >= OVERLAP_MIN_BYTES
    events = None
    if distributed and overlap:
        from ._structs import BWD_GROUPS
        _, events = _overlap_handles(out.device, BWD_GROUPS)
        events = events[:BWD_GROUPS]
        eng.bwd_events, eng.mirror_bf16 = events, grad_dtype == "bf16"
    try:
        loss.backward()
    finally:
        if events is not None:
            eng.bwd_events, eng.mirror_bf16 = None, False
    in_twin = exchange_gradients(eng, events, grad_dtype=grad_dtype) if distributed else False
    if in_twin:
        if getattr(optimizer, "engine", None) is eng and hasattr(optimizer, "step_from_bf16_twin"):
            optimizer.step_from_bf16_twin()        # fused AdamW reads the averaged bf16 bucket (26 B / parameter)
            return loss
        eng.twin_to_flat_grad()                    # any other optimizer: param.grad = fp32(averaged bf16 bucket)
    optimizer.step()
    return loss


class TrainStep:
    """train_step bound to (model, optimizer, loss_fn): `step(data, target, ib, overlap=None)`; info() describes the
    exchange it performs (for bench lines)."""

    def __init__(self, model, optimizer, loss_fn, grad_dtype: str | None = None):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.engine = getattr(model, "_sea_engine", None) or model.engine()
        if grad_dtype is None:
            grad_dtype = "bf16" if getattr(self.engine, "precision", "bf16") == "bf16" else "f32"
        self.grad_dtype = grad_dtype

    def __call__(self, data, target, ib, overlap: bool | None = None):
        return train_step(self.model, self.optimizer, self.loss_fn, data, target, ib, overlap=overlap,
                          grad_dtype=self.grad_dtype)

    def info(self) -> dict:
        eng = self.engine
        buckets = eng.grad_buckets()
        eb = 2 if self.grad_dtype == "bf16" else 4
        plan = plan_buckets(buckets[:-1], eb)
        small = buckets[-1]
        return {"grad_dtype": self.grad_dtype, "buckets": len(plan) + 1, "graphed": False,
                "bucket_bytes": [(e - b) * eb for b, e, _ in plan] + [(small[1] - small[0]) * 4],
                "nccl_bytes_per_step": sum((e - b) * eb for b, e, _ in plan) + (small[1] - small[0]) * 4}

"""Autoregressive rollout through the module's public ``forward`` — the loop of the reference's
``full_autoregressive_evaluation`` / ``autoregressive_validation`` (utils/train_utils.py:202-209,
:170-175): the model is re-run on the whole growing prefix at every step (no KV cache) and only
the last time step of each output is appended."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib
from . import _structs as S


class ProfileSummary(C.Structure):
    _fields_ = [("ms", C.c_double * 3), ("work", C.c_double * 3), ("launches", C.c_int64 * 3)]


def _engine_of(model):
    eng = getattr(model, "_sea_engine", None)
    if eng is None and hasattr(model, "engine"):
        eng = model.engine()
    return eng


def _plan_key(eng):
    """Everything a recorded graph reads through a baked-in pointer: the packed-weight cache, the live
    parameters (``_cache_key[1]``: their data_ptrs — the train/eval flag in ``_cache_key[0]`` is NOT part of
    the key: a plan is always recorded and replayed in inference mode) and the RoPE tables."""
    rope = getattr(eng, "_rope", None)
    return (eng._cache.data_ptr(), eng._cache_key[1],
            None if rope is None else (rope[0].data_ptr(), rope[1].data_ptr()))


def _split_ranges(B: int, splits: int):
    """[(b0, b1)] of at most 4 (the cache holds four stream-K workspaces) near-equal groups of trajectories."""
    splits = max(1, min(int(splits), 4, B))
    base, rem = divmod(B, splits)
    out, b0 = [], 0
    for k in range(splits):
        b1 = b0 + base + (1 if k < rem else 0)
        out.append((b0, b1))
        b0 = b1
    return out


DEFAULT_SPLITS = int(__import__("os").environ.get("SEA_ROLLOUT_SPLITS", "1"))
# consecutive steps recorded into ONE CUDA graph (0 = the whole rollout): a replay is then one cudaGraphLaunch per
# group, and the programmatic-dependent-launch edges between a step's last kernel and the next step's first survive
# (they stop at a graph boundary)
STEPS_PER_GRAPH = int(__import__("os").environ.get("SEA_ROLLOUT_STEPS_PER_GRAPH", "10"))   # measured: 1 -> 5..20: -1 %


def _graph_groups(first: int, last: int, per: int):
    """[(t0, t1)] inclusive step ranges recorded into one CUDA graph each: `per` consecutive steps per graph
    (per <= 0: everything in one graph)."""
    n = last - first + 1
    per = n if per <= 0 else per
    return [(t0, min(last, t0 + per - 1)) for t0 in range(first, last + 1, per)]


class RolloutPlan:
    """The same loop with every step pre-recorded as a CUDA graph.

    Step t (prefix length t) of the reference loop is: forward over the whole prefix, append the
    last output.  Both are recorded once per (B, steps): graph t = [the 22 kernels of
    sea_temporal_forward at T = t, reading the prefix seq[:, :t] in place from the plan's
    [B, steps+1, V, E] sequence buffer] -> [append y[:, -1] to seq[:, t]].  Replaying costs one cudaGraphLaunch per step instead of
    ~30 kernel launches + tensor-map encodes + torch.cat on the host, which is what bounds the short
    prefixes.  Arithmetic, kernels and per-step work are exactly those of the eager loop (the full
    prefix is still recomputed every step); requires a time-invariant ib (checked by the caller).
    STEPS_PER_GRAPH consecutive steps share one graph (scripts/rollout_fuse.py: 33.18 ms at 1, 32.84 at 5 or 20,
    33.28 with the whole rollout in one graph; bit-identical outputs)."""

    def __init__(self, model, B: int, steps: int, device, splits: int = 1):
        eng = _engine_of(model)
        if eng is None:
            raise RuntimeError("RolloutPlan needs a sea_b200 temporal engine (mirror model or accelerate())")
        self.eng, self.B, self.steps = eng, B, steps
        # micro-batches: the B trajectories are independent, and at short prefixes a step is a chain of ~22 small,
        # latency-bound launches that leaves most SMs idle.  `splits` groups of trajectories run the same chain on their
        # own streams inside one graph, so one group's launch / drain latencies are filled with another group's work
        # (same kernels per trajectory: results are bit-identical to the single-batch plan).
        self.subs = _split_ranges(B, splits)
        eng._ensure(False)
        h = eng._h
        V, E, nib = h["V"], h["E"], h["ib_num"]
        self.V, self.E = V, E
        f32 = dict(dtype=torch.float32, device=device)
        self.seq = torch.zeros(B, steps + 1, V, E, **f32)
        self.ib1 = torch.zeros(B, 1, nib, **f32)      # step 1 reads ib as [B,1,nib]
        self.ib2 = torch.zeros(B, 2, nib, **f32)      # step 2 (first time-invariant call) as [B,2,nib]
        self.ws, self.cond, self.ys = [], [], []
        for b0, b1 in self.subs:
            nws = lib.sea_temporal_workspace_bytes(C.byref(eng._desc), b1 - b0, steps, 0)
            self.ws.append(torch.empty(nws, dtype=torch.uint8, device=device))
            ncc = lib.sea_temporal_cond_cache_bytes(C.byref(eng._desc), b1 - b0)
            self.cond.append(torch.empty(ncc, dtype=torch.uint8, device=device))
            self.ys.append(torch.empty((b1 - b0) * steps * V * E, **f32))
        self.streams = [torch.cuda.Stream(device=device) for _ in self.subs[1:]]
        self.graphs, self.launches = [], []
        # the graphs bake in device pointers: packed-weight cache, parameters (biases, norm weights), RoPE tables
        self.key = _plan_key(eng)
        self._rope = eng._rope          # keeps the captured tables alive for the lifetime of the graphs
        self._record()

    def _step(self, t: int) -> int:
        V, E = self.V, self.E
        ib = self.ib1 if t == 1 else self.ib2   # only read while the condition cache is not valid (t <= 2)
        cur = torch.cuda.current_stream()
        for st in self.streams:                  # fork: the other micro-batches start where this step starts
            st.wait_stream(cur)
        n = 0
        for k, (b0, b1) in enumerate(self.subs):
            with torch.cuda.stream(cur if k == 0 else self.streams[k - 1]):
                y = self.ys[k][: (b1 - b0) * t * V * E].view(b1 - b0, t, V, E)
                self.eng._desc.splitk_slot = k
                # the model reads the prefix in place from the sequence buffer (batch-strided view, no gather)
                n += self.eng.forward_into(self.seq[b0:b1, :t], ib[b0:b1], y, self.ws[k], time_invariant=True,
                                           cond_buf=self.cond[k], cond_valid=t > 2)
                self.seq[b0:b1, t].copy_(y[:, t - 1])
        self._step_launches = n + len(self.subs)
        self.eng._desc.splitk_slot = 0
        for st in self.streams:                  # join
            cur.wait_stream(st)
        return n

    def _record(self):
        # one eager pass first: every kernel variant sets its launch attributes outside a capture
        for t in range(1, self.steps + 1):
            self._step(t)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        pool = None
        self.groups = _graph_groups(1, self.steps, STEPS_PER_GRAPH)
        for t0, t1 in self.groups:
            g = torch.cuda.CUDAGraph()
            n = 0
            with torch.cuda.graph(g, pool=pool, stream=side, capture_error_mode="relaxed"):
                for t in range(t0, t1 + 1):
                    n += self._step(t) + len(self.subs)
            pool = g.pool()
            self.graphs.append(g)
            self.launches.append(n)
        self.last_step_launches = self._step_launches
        torch.cuda.current_stream().wait_stream(side)

    def valid_for(self, eng) -> bool:
        return eng._cache is not None and self.key == _plan_key(eng)

    def _host_copy_handles(self, out_host: torch.Tensor):
        """(copy stream, one event per graph group) for streaming the predictions to `out_host`, after checking that it
        is what cudaMemcpy2DAsync can fill asynchronously: pinned fp32 [B, steps, V, E] with contiguous trajectories."""
        want = (self.B, self.steps, self.V, self.E)
        if (out_host.is_cuda or out_host.dtype != torch.float32 or tuple(out_host.shape) != want
                or not out_host.is_pinned() or out_host.stride(3) != 1 or out_host.stride(2) != self.E
                or out_host.stride(1) != self.V * self.E or out_host.stride(0) < self.steps * self.V * self.E):
            raise RuntimeError(f"out_host must be a pinned fp32 host tensor {want} with contiguous trajectories")
        if getattr(self, "_copy", None) is None:
            self._copy = (torch.cuda.Stream(device=self.seq.device), [torch.cuda.Event() for _ in self.graphs])
        return self._copy

    @torch.no_grad()
    def run(self, x0: torch.Tensor, ib: torch.Tensor, out_host: torch.Tensor | None = None) -> torch.Tensor:
        """x0 [B,1,V,E], ib [B,>=1,ib_num] (time-invariant) -> view [B,steps,V,E] of the plan's own
        sequence buffer (overwritten by the next run).  out_host (optional, pinned [B,steps,V,E]): every finished graph
        group's predictions are copied to it on a side stream while the next group runs (one strided
        cudaMemcpy2DAsync per group); the calling stream waits for the copies at the end of the call, so a stream
        synchronize makes out_host readable and a following run cannot overwrite rows still in flight."""
        self.eng._ensure(False)
        self.seq[:, 0].copy_(x0[:, 0])
        self.ib1.copy_(ib[:, :1])
        self.ib2.copy_(ib[:, :1].expand(-1, 2, -1))
        if out_host is None:
            for g in self.graphs:
                g.replay()
        else:
            copy_stream, events = self._host_copy_handles(out_host)
            cur = torch.cuda.current_stream()
            row = self.V * self.E * 4                           # bytes of one step of one trajectory
            cs = C.c_void_p(copy_stream.cuda_stream)
            for g, ev, (t0, t1) in zip(self.graphs, events, self.groups):
                g.replay()
                ev.record(cur)
                copy_stream.wait_event(ev)
                # seq[:, t0 : t1+1] (steps t0..t1 of every trajectory) -> out_host[:, t0-1 : t1]
                check(lib.sea_copy_rows_to_host(C.c_void_p(out_host.data_ptr() + (t0 - 1) * row),
                                                C.c_size_t(out_host.stride(0) * 4),
                                                C.c_void_p(self.seq.data_ptr() + t0 * row),
                                                C.c_size_t(self.seq.stride(0) * 4),
                                                C.c_size_t((t1 - t0 + 1) * row), C.c_size_t(self.B), cs),
                      "copy_rows_to_host")
            cur.wait_stream(copy_stream)
        n = sum(self.launches)
        self.eng.last_launches = self.last_step_launches
        self.eng.total_launches += n
        return self.seq[:, 1:]


class CachedRolloutPlan:
    """KV-cached rollout: every step feeds only the newest token of each
    trajectory through ``sea_temporal_step``; keys / values of earlier positions come from a cache, so
    a step costs O(1) forward work instead of the reference loop's O(t) prefix recompute.  The model is
    causal, hence the outputs equal the prefix loop's up to rounding (different reduction order in the
    attention).  One CUDA graph per step (the cached length is baked into each launch); the step reads
    x_t from and writes y_t into one [B, steps+1, V, E] sequence buffer, so a step has no copy kernels.
    Works for time-invariant and time-varying ib."""

    def __init__(self, model, B: int, steps: int, device, time_invariant: bool, splits: int = 1):
        eng = _engine_of(model)
        if eng is None:
            raise RuntimeError("CachedRolloutPlan needs a sea_b200 temporal engine")
        self.eng, self.B, self.steps, self.inv = eng, B, steps, bool(time_invariant)
        self.subs = _split_ranges(B, splits)     # micro-batches on their own streams (see RolloutPlan)
        eng._ensure(False)
        h = eng._h
        if h["src_len"] != 0:
            raise NotImplementedError("cached rollout needs src_len == 0: with tril(diagonal=src_len>0) a query "
                                      "sees future keys and the prefix loop is not reproducible from a KV cache")
        V, E, nib = h["V"], h["E"], h["ib_num"]
        f32 = dict(dtype=torch.float32, device=device)
        self.seq = torch.zeros(B, steps + 1, V, E, **f32)
        self.ib = torch.zeros(B, steps, nib, **f32)
        d = C.byref(eng._desc)
        u8 = dict(dtype=torch.uint8, device=device)
        self.kv = [torch.empty(lib.sea_temporal_kv_cache_bytes(d, b1 - b0, steps), **u8) for b0, b1 in self.subs]
        self.ws = [torch.empty(lib.sea_temporal_workspace_bytes(d, b1 - b0, 1, 0), **u8) for b0, b1 in self.subs]
        self.cond = [torch.empty(lib.sea_temporal_cond_cache_bytes(d, b1 - b0), **u8) for b0, b1 in self.subs]
        self.streams = [torch.cuda.Stream(device=device) for _ in self.subs[1:]]
        self.graphs, self.launches = [], []
        # the graphs bake in device pointers: packed-weight cache, parameters (biases, norm weights), RoPE tables
        self.key = _plan_key(eng)
        self._rope = eng._rope          # keeps the captured tables alive for the lifetime of the graphs
        self._record()

    def _step(self, t: int) -> int:
        cur = torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(cur)
        n = 0
        for k, (b0, b1) in enumerate(self.subs):
            with torch.cuda.stream(cur if k == 0 else self.streams[k - 1]):
                self.eng._desc.splitk_slot = k
                n += self.eng.step_into(self.seq[b0:b1, t], self.ib[b0:b1, t], self.seq[b0:b1, t + 1], t, self.kv[k],
                                        self.steps, self.ws[k], time_invariant=self.inv, cond_buf=self.cond[k],
                                        cond_valid=t > 0)
        self.eng._desc.splitk_slot = 0
        for st in self.streams:
            cur.wait_stream(st)
        return n

    def _record(self):
        for t in range(self.steps):      # eager pass: launch attributes are set outside any capture
            self._step(t)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        pool = None
        for t0, t1 in _graph_groups(0, self.steps - 1, STEPS_PER_GRAPH):
            g = torch.cuda.CUDAGraph()
            n = 0
            with torch.cuda.graph(g, pool=pool, stream=side, capture_error_mode="relaxed"):
                for t in range(t0, t1 + 1):
                    self.last_step_launches = self._step(t)
                    n += self.last_step_launches
            pool = g.pool()
            self.graphs.append(g)
            self.launches.append(n)
        torch.cuda.current_stream().wait_stream(side)

    def valid_for(self, eng) -> bool:
        return eng._cache is not None and self.key == _plan_key(eng)

    @torch.no_grad()
    def run(self, x0: torch.Tensor, ib: torch.Tensor) -> torch.Tensor:
        self.eng._ensure(False)
        self.seq[:, 0].copy_(x0[:, 0])
        self.ib.copy_(ib[:, : self.steps])
        for g in self.graphs:
            g.replay()
        self.eng.last_launches = self.last_step_launches
        self.eng.total_launches += sum(self.launches)
        return self.seq[:, 1:]


def _profiling() -> bool:
    return bool(getattr(profile, "active", False))


@torch.no_grad()
def rollout(model, x0: torch.Tensor, ib: torch.Tensor, steps: int,
            ib_time_invariant: bool | None = None, graphs: bool = True, cached: bool = False,
            _view_ok: bool = False, splits: int | None = None, out_host: torch.Tensor | None = None) -> torch.Tensor:
    """x0 [B,1,V,E], ib [B,>=steps,ib_num] -> predicted latents [B,steps,V,E].

    ``ib`` is the time-invariant physical parameter of a trajectory in the reference's data
    (models/temporal.py:111-120 "TIPI").  When every ib[b, t] equals ib[b, 0] (checked once here on
    the device unless the caller passes the answer), the AdaLN cond_mlp and the TIPI MLP are
    evaluated once per trajectory instead of once per token; the result is the same function of
    the same inputs.

    ``cached=True`` (opt-in) runs the KV-cached incremental engine instead of recomputing the prefix:
    same outputs up to rounding, O(1) instead of O(t) work per step (``CachedRolloutPlan``).

    ``out_host`` (optional, pinned fp32 [B,steps,V,E]): also deliver the predictions to the host, asynchronously on the
    calling stream's timeline (synchronize the stream before reading it).  The graphed plan streams every finished
    group of steps out while the next one runs; the other paths copy once at the end."""
    eng = _engine_of(model)
    if ib_time_invariant is None:
        ib_time_invariant = bool((ib[:, :steps] == ib[:, :1]).all().item())
    if cached:
        if eng is None or not x0.is_cuda:
            raise RuntimeError("cached rollout needs a sea_b200 temporal engine on a CUDA device")
        eng._ensure(False)
        plans = eng.__dict__.setdefault("_cached_plans", {})
        nsplit = DEFAULT_SPLITS if splits is None else int(splits)
        key = (x0.shape[0], steps, x0.device.index, bool(ib_time_invariant), nsplit)
        plan = plans.get(key)
        if plan is None or not plan.valid_for(eng):
            if len(plans) >= 4:
                plans.clear()
            plan = plans[key] = CachedRolloutPlan(model, x0.shape[0], steps, x0.device, bool(ib_time_invariant), nsplit)
        out = plan.run(x0, ib)
        if out_host is not None:
            out_host.copy_(out, non_blocking=True)
        return out if _view_ok else out.clone()
    if graphs and eng is not None and ib_time_invariant and x0.is_cuda and not _profiling():
        eng._ensure(False)
        plans = eng.__dict__.setdefault("_rollout_plans", {})
        nsplit = DEFAULT_SPLITS if splits is None else int(splits)
        key = (x0.shape[0], steps, x0.device.index, nsplit)
        plan = plans.get(key)
        if plan is None or not plan.valid_for(eng):
            if len(plans) >= 4:
                plans.clear()
            plan = plans[key] = RolloutPlan(model, x0.shape[0], steps, x0.device, nsplit)
        out = plan.run(x0, ib, out_host=out_host)
        return out if _view_ok else out.clone()   # the plan's buffer is overwritten by the next run
    prev = None
    if eng is not None:
        prev, eng.ib_time_invariant = eng.ib_time_invariant, ib_time_invariant
        # same trajectories, same weights, same ib for every step of this loop
        eng.cond_reuse, eng._cond_valid = bool(ib_time_invariant), False
    try:
        seq = x0
        for i in range(steps):
            out = model(seq, ib[:, : i + 1])
            seq = torch.cat((seq, out[:, -1:]), dim=1)
            if eng is not None:
                eng.weights_frozen = True   # checked once by the first call of this loop
    finally:
        if eng is not None:
            eng.ib_time_invariant = prev
            eng.cond_reuse, eng._cond_valid = False, False
            eng.weights_frozen = False
    if out_host is not None:
        out_host.copy_(seq[:, 1:], non_blocking=True)
    return seq[:, 1:]


def rollout_from_host(model, x0_host: torch.Tensor, ib_host: torch.Tensor, steps: int,
                      out_host: torch.Tensor, device) -> torch.Tensor:
    """End-to-end variant: pinned host inputs -> device, rollout, predicted latents -> pinned host (streamed out group by
    group behind the graphed plan, see ``RolloutPlan.run``)."""
    x0 = x0_host.to(device, non_blocking=True)
    ib = ib_host.to(device, non_blocking=True)
    rollout(model, x0, ib, steps, _view_ok=True, out_host=out_host)
    torch.cuda.current_stream().synchronize()
    return out_host


class profile:
    """Context manager around lib.sea_profile_begin/end; ``.summary`` holds per-category totals."""
    CATS = ("gemm", "attention", "elementwise")

    active = False

    def __enter__(self):
        lib.sea_profile_begin()
        profile.active = True
        return self

    def __exit__(self, *exc):
        profile.active = False
        s = ProfileSummary()
        check(lib.sea_profile_end(C.byref(s)), "profile_end")
        self.summary = {c: dict(ms=s.ms[i], work=s.work[i], launches=int(s.launches[i]))
                        for i, c in enumerate(self.CATS)}
        return False

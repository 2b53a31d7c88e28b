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


@torch.no_grad()
def rollout(model, x0: torch.Tensor, ib: torch.Tensor, steps: int,
            ib_time_invariant: bool | None = None) -> torch.Tensor:
    """x0 [B,1,V,E], ib [B,>=steps,ib_num] -> predicted latents [B,steps,V,E].

    ``ib`` is the time-invariant physical parameter of a trajectory in the reference's data
    (models/temporal.py:111-120 "TIPI").  When every ib[b, t] equals ib[b, 0] (checked once here on
    the device unless the caller passes the answer), the AdaLN cond_mlp and the TIPI MLP are
    evaluated once per trajectory instead of once per token; the result is the same function of
    the same inputs."""
    eng = _engine_of(model)
    if ib_time_invariant is None:
        ib_time_invariant = bool((ib[:, :steps] == ib[:, :1]).all().item())
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
    return seq[:, 1:]


def rollout_from_host(model, x0_host: torch.Tensor, ib_host: torch.Tensor, steps: int,
                      out_host: torch.Tensor, device) -> torch.Tensor:
    """End-to-end variant: pinned host inputs -> device, rollout, predicted latents -> pinned host."""
    x0 = x0_host.to(device, non_blocking=True)
    ib = ib_host.to(device, non_blocking=True)
    pred = rollout(model, x0, ib, steps)
    out_host.copy_(pred, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host


class profile:
    """Context manager around lib.sea_profile_begin/end; ``.summary`` holds per-category totals."""
    CATS = ("gemm", "attention", "elementwise")

    def __enter__(self):
        lib.sea_profile_begin()
        return self

    def __exit__(self, *exc):
        s = ProfileSummary()
        check(lib.sea_profile_end(C.byref(s)), "profile_end")
        self.summary = {c: dict(ms=s.ms[i], work=s.work[i], launches=int(s.launches[i]))
                        for i, c in enumerate(self.CATS)}
        return False

"""Fused AdamW for the temporal model.

Drop-in for the optimizer the reference builds in ``utils/train_utils.py:33-39``
(``torch.optim.AdamW(model.parameters(), lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)``) and
steps at ``train/train_temporal.py:258``: same constructor arguments, same ``param_groups`` /
``state`` layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter, so ``state_dict()`` moves
both ways), same arithmetic — but ``step()`` is ONE kernel launch (``sea_adamw_step``) over a
device-resident chunk table instead of torch's multi-pass foreach implementation, and it writes
the refreshed bf16 tensor-core copies of the weights directly into the engine's packed cache, so
the next forward only has to rebuild the transposes used by dgrad.

There is no CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from ._lib import check, lib

CHUNK = 16384  # elements per CTA

_CHUNK_DT = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("b", "<u8"),
                      ("n", "<i4"), ("r", "<i4")])


class AdamWHyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("bias_corr1", C.c_float), ("bias_corr2_sqrt", C.c_float),
                ("grad_scale", C.c_float), ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float)]


class AdamW(torch.optim.Optimizer):
    """``AdamW(params, lr, betas, eps, weight_decay)`` — torch.optim.AdamW semantics (no amsgrad,
    no maximize).  ``engine`` (optional, a ``TemporalEngine``): lets ``step`` write the bf16 weight
    copies in the same pass (an engine that is NOT passed here still notices the update through the
    parameters' version counters and re-packs its copies on its next call); ``grad_scale`` folds e.g. the data-parallel 1/world into the step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *,
                 engine=None, grad_scale: float = 1.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.engine = engine
        self.grad_scale = float(grad_scale)
        self._tables = {}
        self._steps = {}     # group index -> [ids of its stepped params, host step count, shared step tensor]

    def _table(self, gi, plist, twin=False):
        """Device chunk table of one param group, rebuilt only when a pointer changes.  twin: weight gradients are read
        from the engine's averaged bf16 bucket (data-parallel exchange, sea_b200.parallel) instead of param.grad."""
        eng = self.engine
        gi = (gi, bool(twin))
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in plist)
        if twin:
            flat, tw, n_w = eng.flat_grad(), eng.flat_grad_bf16(), eng.grad_buckets()[-1][0]
            key = key + (flat.data_ptr(), tw.data_ptr())
        ckey = None if eng is None else ((eng._cache.data_ptr() if eng._cache is not None else 0),
                                         eng._cache_key[0] if eng._cache_key else None)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key and hit[1] == ckey:
            return hit[2], hit[3]
        rows = []
        fused_copies = False
        for p in plist:
            st = self.state[p]
            slot = 0
            if eng is not None and eng._cache is not None and eng._desc is not None and eng.precision == "bf16":
                out = C.c_void_p()
                check(lib.sea_temporal_cache_slot(C.byref(eng._desc), C.c_void_p(eng._cache.data_ptr()),
                                                  int(eng._cache_key[0]) if eng._cache_key else 1,
                                                  C.c_void_p(p.data_ptr()), C.byref(out)), "cache_slot")
                slot = out.value or 0
                fused_copies |= slot != 0
            n = p.numel()
            g_ptr, g_sz, g_b16 = p.grad.data_ptr(), 4, 0
            if twin:
                e0 = (g_ptr - flat.data_ptr()) // 4
                if 0 <= e0 < n_w:                     # a weight-gradient slot of the flat buffer: its bf16 twin
                    g_ptr, g_sz, g_b16 = tw.data_ptr() + 2 * e0, 2, 1
            for off in range(0, n, CHUNK):
                c = min(CHUNK, n - off)
                rows.append((p.data_ptr() + 4 * off, g_ptr + g_sz * off,
                             st["exp_avg"].data_ptr() + 4 * off, st["exp_avg_sq"].data_ptr() + 4 * off,
                             (slot + 2 * off) if slot else 0, c, g_b16))
        arr = np.array(rows, dtype=_CHUNK_DT)
        dev = plist[0].device
        tab = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self._tables[gi] = (key, ckey, tab, fused_copies)
        return tab, fused_copies

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._steps = {}
        self._tables = {}

    def step_from_bf16_twin(self):
        """step() with the weight gradients taken from the engine's averaged bf16 bucket (what
        sea_b200.parallel.exchange_gradients(..., grad_dtype="bf16") leaves behind): 26 B / parameter instead of 28."""
        if self.engine is None:
            raise RuntimeError("step_from_bf16_twin needs the engine that owns the gradient buckets")
        return self.step(_twin=True)

    @torch.no_grad()
    def step(self, closure=None, _twin=False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        wrote_copies = False
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("sea_b200.optim.AdamW: parameters must be contiguous fp32 CUDA tensors")
                if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    raise RuntimeError("sea_b200.optim.AdamW: gradients must be contiguous fp32")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            # one step count per group: every parameter's `step` entry (torch's state layout) is the SAME
            # tensor object, so a step costs one tiny update instead of one per parameter
            key = tuple(id(p) for p in plist)
            track = self._steps.get(gi)
            if track is None or track[0] != key or any(self.state[p]["step"] is not track[2] for p in plist[:1]):
                steps = {float(self.state[p]["step"]) for p in plist}
                if len(steps) != 1:
                    raise RuntimeError("sea_b200.optim.AdamW: parameters of one group must share the step count")
                shared = torch.tensor(steps.pop(), dtype=torch.float32)
                for p in plist:
                    self.state[p]["step"] = shared
                track = [key, float(shared), shared]
                self._steps[gi] = track
            track[1] += 1.0
            track[2] += 1.0
            t = track[1]
            b1, b2 = group["betas"]
            hp = AdamWHyper(group["lr"], b1, b2, group["eps"], group["weight_decay"],
                            1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), self.grad_scale, 1.0 - b1, 1.0 - b2)
            tab, fused = self._table(gi, plist, _twin)
            with torch.cuda.device(plist[0].device):
                check(lib.sea_adamw_step(C.c_void_p(tab.data_ptr()), tab.numel() // _CHUNK_DT.itemsize,
                                         C.byref(hp), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "adamw_step")
            wrote_copies |= fused
            # the kernel wrote the masters through raw pointers: tell autograd / every version-keyed cache (any
            # TemporalEngine on these parameters, not only `self.engine`) that they changed
            torch._C._increment_version(plist)
        if self.engine is not None:
            self.engine.after_optimizer_step(straight_copies_fresh=wrote_copies)
        return loss

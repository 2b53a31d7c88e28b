"""Autograd bridge for the temporal engine.

The whole model is ONE autograd node: forward = one FFI call that leaves its tape in a workspace,
backward = one FFI call that reads it.  Parameter gradients are accumulated by the CUDA kernels
straight into ``param.grad`` (views of one flat fp32 buffer owned by the engine — the same buffer
the data-parallel all-reduce runs on), with torch's own semantics: ``grad is None`` -> fresh zeros,
otherwise ``+=``; the reference's dead parameters keep ``grad = None``.

Consequences of writing ``param.grad`` from inside the node (documented limits): ``torch.autograd.grad`` w.r.t.
parameters and parameter hooks do not see these gradients (use ``loss.backward()`` as the reference loop does,
train/train_temporal.py:257); the tape is released by the first backward, so a second backward over the same
graph raises instead of reading a recycled workspace.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib


class _TemporalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ib, anchor, engine):
        B, T = x.shape[0], x.shape[1]
        ws = engine.acquire_training_workspace(B, T)
        y = engine.forward_nograd(x, ib, training=True, ws=ws)
        ctx.engine, ctx.ws, ctx.shape = engine, ws, (B, T)
        ctx.dropout = (engine.last_dropout_seed, float(engine.dropout))   # the backward regenerates the masks
        ctx.save_for_backward(x, ib)
        ctx.need_dx = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.engine
        if ctx.ws is None:
            raise RuntimeError("sea_b200: the activation tape of this forward was released by its first backward "
                               "(retain_graph / double backward are not supported; run the forward again)")
        x, ib = ctx.saved_tensors
        dx = eng.backward(x, ib, dy, ctx.ws, ctx.need_dx, dropout_seed=ctx.dropout[0], dropout_p=ctx.dropout[1])
        eng.release_training_workspace(ctx.shape, ctx.ws)
        ctx.ws = None
        return dx, None, None, None


def temporal_apply(engine, x, ib):
    anchor = engine.anchor_param()
    return _TemporalFn.apply(x, ib, anchor, engine)

"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY.

Deterministic, machine-independent recipe for the weights and inputs of the golden cases.
Every tensor is drawn from numpy's frozen legacy ``RandomState`` stream (NEP 19 guarantees the
stream), seeded per NAME, so the build container (which has the reference) and the GPU box
(which does not) regenerate bit-identical weights without shipping them; only the reference's
OUTPUTS are committed under tests/golden/.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

SKIP_SUFFIXES = (".tril", ".freqs_cis", ".pe")  # buffers of the reference, not parameters


def _rs(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def tensor_for(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """Value of parameter ``name``: Linear weights ~ N(0, 0.02) as in the reference's
    _init_weights (models/temporal.py:395-402); biases and norm affine parameters get small
    non-zero values so that every term of every kernel is exercised."""
    r = _rs(name, seed).standard_normal(shape).astype(np.float32)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = any(t in name for t in (".ln.", "ln_cross", "ln_exp", ".layers.1.", "encode.ln")) \
        and ".cond_mlp." not in name
    if name.startswith("ln.") and ".cond_mlp." not in name:
        is_norm = True
    if is_norm:
        val = 1.0 + 0.05 * r if leaf == "weight" else 0.05 * r
    elif leaf == "weight":
        fan_in = shape[-1]
        # patch MLPs / decoder keep torch-default-like scale; everything else N(0, 0.02)
        val = r * (1.0 / np.sqrt(3 * fan_in)) if ("encoders." in name or "decoders." in name) else 0.02 * r
        if ".cond_mlp.0." in name or ".ib.layers.0." in name:
            val = 0.5 * r  # scalar-input layers: make the ib dependence visible
    else:
        val = 0.01 * r
    return torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32))


def fill_state(shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int) -> Dict[str, torch.Tensor]:
    return {n: tensor_for(n, tuple(s), seed) for n, s in shapes
            if not n.endswith(SKIP_SUFFIXES)}


def temporal_inputs(B: int, T: int, V: int, E: int, seed: int, ib_num: int = 1):
    x = torch.from_numpy(_rs("x", seed).standard_normal((B, T, V, E)).astype(np.float32))
    ib = torch.from_numpy(_rs("ib", seed).uniform(0, 1, (B, T, ib_num)).astype(np.float32))
    tgt = torch.from_numpy(_rs("target", seed).standard_normal((B, T, V, E)).astype(np.float32))
    return x, ib, tgt


def spatial_inputs(B: int, P: int, Fn: int, C: int, seed: int):
    r = _rs("fields", seed)
    x = r.standard_normal((B, P, Fn, C)).astype(np.float32)
    keep = r.randint(C // 2, C + 1, size=(B, P, 1, 1))           # zero-padded tail per patch
    x = x * (np.arange(C).reshape(1, 1, 1, C) < keep)
    x[0, 0, 0, -1] = -9999.0                                      # exercise generate_padding_mask
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def probe_vector(name: str, shape, seed: int) -> torch.Tensor:
    return torch.from_numpy(_rs("probe:" + name, seed).standard_normal(tuple(shape)).astype(np.float32))


def temporal_shapes(*, embed_dim, n_heads, scale_ratio, num_variables, down_proj=2, num_layers=1,
                    ln_type="adaln", ib_num=1):
    """(name, shape) of every LIVE reference parameter, reference naming."""
    del n_heads
    return list(_temporal_shapes_lazy(embed_dim, scale_ratio, num_variables, down_proj, num_layers,
                                      ln_type, ib_num))


def _temporal_shapes_lazy(E, scale_ratio, V, down_proj, L, ln_type, ib_num):
    Dd, H = E // down_proj, int(E * scale_ratio)

    def lin(n, o, i, bias=True):
        yield n + ".weight", (o, i)
        if bias:
            yield n + ".bias", (o,)

    def nrm(n, d, kind, with_bias):
        yield n + ".weight", (d,)
        if kind == "adaln":
            yield n + ".bias", (d,)
            yield from lin(n + ".cond_mlp.0", 2 * d, ib_num)
            yield from lin(n + ".cond_mlp.2", 2 * d, 2 * d)
        elif with_bias:
            yield n + ".bias", (d,)

    for layer in range(L):
        p = f"blocks.{layer}"
        for i in range(V):
            yield from nrm(f"{p}.ln.exp.{i}.0", E, ln_type, False)
            yield from nrm(f"{p}.ln.exp.{i}.2", E, ln_type, False)
            for nm in ("q", "k", "v"):
                yield from lin(f"{p}.attn.self.{i}.{nm}", E, E)
            yield from lin(f"{p}.attn.self.{i}.projection", E, E, False)
            yield from lin(f"{p}.cross_down.{i}", Dd, E)
            yield from lin(f"{p}.cross_up.{i}", E, Dd)
            yield from nrm(f"{p}.ln_cross.{i}", Dd, ln_type, False)
            for j in range(V):
                if j != i:
                    for nm in ("q", "k", "v"):
                        yield from lin(f"{p}.cross_attn.{i}.{j}.{nm}", Dd, Dd)
                    yield from lin(f"{p}.cross_attn.{i}.{j}.projection", Dd, Dd, False)
            yield from lin(f"{p}.mlp.{i}.layers.0", H, E)
            yield from nrm(f"{p}.mlp.{i}.layers.1", H, "ln", True)
            yield from lin(f"{p}.mlp.{i}.layers.3", E, H)
            yield from lin(f"{p}.proj.{i}", E, E)
        hid = max(1, int(ib_num * scale_ratio))
        yield from lin(f"{p}.ib.layers.0", hid, ib_num)
        yield from nrm(f"{p}.ib.layers.1", hid, "ln", True)
        yield from lin(f"{p}.ib.layers.3", E, hid)
    for i in range(V):
        yield from nrm(f"ln.{i}", E, ln_type, False)

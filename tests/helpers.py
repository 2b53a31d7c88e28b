"""Shared test helpers: golden loading + oracle-side state construction (checker only)."""
import os

import numpy as np
import torch

from oracle import golden_recipe as gr
from oracle import sea_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 20241018


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def temporal_case(tag, ln_type, requires_grad=False, dtype=torch.float32):
    g = load_golden("temporal_" + tag)
    E, nh, scale, V, B, T, steps = [int(v) for v in g["meta"]]
    shapes = gr.temporal_shapes(embed_dim=E, n_heads=nh, scale_ratio=scale, num_variables=V,
                                ln_type=ln_type)
    sd = {k: v.to(dtype) for k, v in gr.fill_state(shapes, SEED).items()}
    if requires_grad:
        for v in sd.values():
            v.requires_grad_(True)
    x, ib, tgt = gr.temporal_inputs(B, T, V, E, SEED)
    cfg = dict(num_layers=1, n_heads=nh, ln_type=ln_type)
    return g, sd, cfg, x.to(dtype), ib.to(dtype), tgt.to(dtype), steps


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


TEMPORAL_CASES = [("small_adaln", "adaln"), ("small_ln", "ln"), ("small_v3", "ln"),
                  ("cylinder_flow", "adaln"), ("multiphase_flow", "ln")]
SPATIAL_CASES = {"small": dict(n_inp=16, mlp_hidden=48, num_layers=2, embed_dim=8, n_heads=8),
                 "cylinder_flow": dict(n_inp=64, **{k: v for k, v in so.SPATIAL_CONFIGS["cylinder_flow"].items() if k != "field_groups"}),
                 "multiphase_flow": dict(n_inp=64, **{k: v for k, v in so.SPATIAL_CONFIGS["multiphase_flow"].items() if k != "field_groups"})}


def spatial_case(tag):
    g = load_golden("spatial_" + tag)
    c = SPATIAL_CASES[tag]
    n_inp, hidden, layers, D, nh, B = [int(v) for v in g["meta"]]
    fg = [[0, 1], [2]]
    shapes = [(k, tuple(v.shape)) for k, v in so.init_spatial_state(
        field_groups=fg, n_inp=n_inp, mlp_hidden=hidden, num_layers=layers, embed_dim=D).items()]
    sd = gr.fill_state(shapes, SEED)
    x = gr.spatial_inputs(B, 64, 3, n_inp, SEED)
    return g, sd, dict(field_groups=fg, num_layers=layers, n_heads=nh), x

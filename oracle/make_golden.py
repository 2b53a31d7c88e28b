"""ORACLE SUPPORT — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden.py

For every case the reference's own nn.Module classes are constructed with the reference's own
constructor arguments, their parameters are overwritten with the machine-independent recipe of
oracle/golden_recipe.py, and the reference forward / backward / rollout outputs are stored.
Weights and inputs are NOT stored: tests regenerate them from the same recipe.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
for stub in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "h5py", "wandb"):
    sys.modules.setdefault(stub, types.ModuleType(stub))

from oracle import golden_recipe as gr  # noqa: E402

from models.temporal import TemporalModel  # noqa: E402  (reference)
from models.encoder_decoder import SpatialModel  # noqa: E402  (reference)
from utils.train_utils import transform_processed_data, inverse_transform_processed_data  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 20241018


def load_recipe(model, seed):
    sd = model.state_dict()
    new = {}
    for name, t in sd.items():
        if name.endswith(gr.SKIP_SUFFIXES):
            continue
        new[name] = gr.tensor_for(name, tuple(t.shape), seed)
    missing = model.load_state_dict(new, strict=False)
    assert all(k.endswith(gr.SKIP_SUFFIXES) for k in missing.missing_keys), missing
    return model


def temporal_case(tag, *, E, nh, scale, V, ln, B, T, max_len=64, rollout_steps=0, grads=True):
    torch.manual_seed(0)
    m = TemporalModel(1, E, nh, max_len, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1,
                      True, ln)
    load_recipe(m, SEED)
    m.eval()
    x, ib, tgt = gr.temporal_inputs(B, T, V, E, SEED)
    out = {"meta": np.array([E, nh, scale, V, B, T, rollout_steps], dtype=np.int64)}
    y = m(x, ib)
    out["y"] = y.detach().numpy()
    if grads:
        loss = torch.nn.functional.mse_loss(y, tgt)
        loss.backward()
        out["loss"] = np.array(loss.item(), dtype=np.float64)
        names, norms, probes = [], [], []
        for name, p in m.named_parameters():
            if p.grad is None:
                continue
            names.append(name)
            norms.append(p.grad.double().norm().item())
            probes.append((p.grad.double() * gr.probe_vector(name, p.shape, SEED).double()).sum().item())
            if p.grad.numel() <= 4096:
                out["grad:" + name] = p.grad.numpy().copy()
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms)
        out["grad_probes"] = np.array(probes)
        dead = [n for n, p in m.named_parameters() if p.grad is None]
        out["dead_params"] = np.array(dead)
    if rollout_steps:
        with torch.no_grad():
            seq = x[:1, :1]
            ibr = ib[:1].repeat(1, (rollout_steps + T - 1) // T + 1, 1)[:, :rollout_steps]
            for i in range(rollout_steps):  # utils/train_utils.py:203-207
                o = m(seq, ibr[:, : i + 1])
                seq = torch.cat((seq, o[:, -1:]), dim=1)
            out["rollout"] = seq[:, 1:].numpy()
    np.savez_compressed(os.path.join(OUT, f"temporal_{tag}.npz"), **out)
    print(tag, "y", out["y"].shape, "loss", out.get("loss"))


def spatial_case(tag, *, n_inp, hidden, layers, D, nh, B):
    torch.manual_seed(0)
    fg = [[0, 1], [2]]
    m = SpatialModel(fg, n_inp, hidden, layers, D, nh, 2024, 0, 0.0, False)
    load_recipe(m, SEED)
    m.eval()
    x = gr.spatial_inputs(B, 64, 3, n_inp, SEED)
    with torch.no_grad():
        xin = x.clone()
        y = m(xin)                      # mutates xin in place (pad rewrite)
        z = m.encode(xin)
        lat = transform_processed_data(z, 1, B, 64, 2)
        back = inverse_transform_processed_data(lat, 1, B, 64, 2)
    assert torch.equal(back, z)
    np.savez_compressed(os.path.join(OUT, f"spatial_{tag}.npz"),
                        meta=np.array([n_inp, hidden, layers, D, nh, B], dtype=np.int64),
                        y=y.numpy(), z=z.numpy(), latent=lat.numpy(),
                        x_after=xin.numpy()[0, 0, 0, -4:])
    print(tag, "y", tuple(y.shape), "z", tuple(z.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # small cases: every code path, every parameter gradient
    temporal_case("small_adaln", E=128, nh=2, scale=2, V=2, ln="adaln", B=2, T=12, rollout_steps=6)
    temporal_case("small_ln", E=128, nh=2, scale=2, V=2, ln="ln", B=3, T=9, rollout_steps=5)
    temporal_case("small_v3", E=128, nh=2, scale=2, V=3, ln="ln", B=1, T=7)
    # the two real configs (configs/cylinder_flow.py, configs/multiphase_flow.py): forward on a
    # short window + the 10-step batch-1 rollout of BASELINE.json configs[0]
    temporal_case("cylinder_flow", E=1024, nh=8, scale=8, V=2, ln="adaln", B=1, T=16,
                  max_len=64, rollout_steps=10, grads=True)
    temporal_case("multiphase_flow", E=2048, nh=8, scale=8, V=2, ln="ln", B=1, T=16,
                  max_len=64, rollout_steps=10, grads=True)
    spatial_case("small", n_inp=16, hidden=48, layers=2, D=8, nh=8, B=3)
    spatial_case("cylinder_flow", n_inp=64, hidden=480, layers=12, D=16, nh=8, B=2)
    spatial_case("multiphase_flow", n_inp=64, hidden=624, layers=12, D=32, nh=8, B=2)


def state_dict_manifest():
    """Names/shapes/dtypes of the reference state_dict (checkpoint-compatibility contract)."""
    import json
    out = {}
    for tag, (E, nh, scale, V, ln) in {"small_adaln": (128, 2, 2, 2, "adaln"), "small_ln": (128, 2, 2, 2, "ln"),
                                       "small_v3": (128, 2, 2, 3, "ln")}.items():
        m = TemporalModel(1, E, nh, 64, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
        out["temporal_" + tag] = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    m = SpatialModel([[0, 1], [2]], 16, 48, 2, 8, 8, 2024, 0, 0.0, False)
    out["spatial_small"] = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_manifest.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    state_dict_manifest()

"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not imported by the product path.

A functional CPU restatement (torch fp32/fp64 tensor arithmetic, no nn.Module, no CUDA) of the
reference's hot path, written from the reference's behaviour:

    models/temporal.py       TemporalModel.forward :405-416, BaseBlockTemporal.forward :126-148,
                             SEABlockTemporal._apply_cross_attention / _apply_exchange :176-192,
                             _add_info :111-120
    models/base_blocks.py    MLP :9-47, up/downScaleMLP :49-78, LayerNorm :80-88,
                             MultiHeadAttention :91-121, EncoderBlock :123-138,
                             MaskedMultiHeadAttention :155-203, MaskedMultiHeadCrossAttention :246-295,
                             RoPE :300-324, AdaLN :330-350, PositionalEncoding :355-372
    models/encoder_decoder.py PointwiseEncode.forward :105-123, Decode.forward :137-146,
                             SpatialModel.forward :161-176

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this file, and only as the checker / the timed CPU baseline.

Parity pin: the reference ships NO golden vectors or known-answer tests, so the
oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the
unchanged reference modules from /root/reference (build container only), runs them on seeded
weights/inputs and commits inputs + outputs + gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those fixtures on every CPU run.

Parameters are addressed by the reference's own ``state_dict()`` names, so a reference checkpoint
feeds the oracle unchanged.  Everything is differentiable (torch autograd on CPU) so the same
functions give reference gradients for the backward kernels.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------- primitives
def linear(x, sd: State, name: str):
    """nn.Linear: y = x W^T + b (bias optional)."""
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def gelu(x):
    """nn.GELU() default = exact erf form (models/base_blocks.py:25,56,71)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def layer_norm(x, weight, bias=None, eps: float = 1e-5):
    """F.layer_norm over the last dim, biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps) * weight
    return y if bias is None else y + bias


def adaln(x, cond, sd: State, name: str):
    """AdaLN.forward, models/base_blocks.py:343-350."""
    c = F.linear(cond, sd[name + ".cond_mlp.0.weight"], sd[name + ".cond_mlp.0.bias"])
    c = c * torch.sigmoid(c)  # SiLU
    c = F.linear(c, sd[name + ".cond_mlp.2.weight"], sd[name + ".cond_mlp.2.bias"])
    w, b = c.chunk(2, dim=-1)
    w = w + 1
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    xn = (x - mu) / torch.sqrt(var + 1e-5)
    return xn * (sd[name + ".weight"] + w) + (sd[name + ".bias"] + b)


def norm(x, cond, sd: State, name: str, ln_type: str):
    """LN_type switch of models/temporal.py:63-74 / :165-172 / :367-372."""
    if ln_type == "adaln":
        return adaln(x, cond, sd, name)
    return layer_norm(x, sd[name + ".weight"], sd.get(name + ".bias"))


def rope_table(head_dim: int, length: int, theta: float = 10000.0, dtype=torch.float32):
    """precompute_freqs_cis, models/base_blocks.py:300-305 → (cos, sin) of shape [length, hd/2]."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(length, dtype=torch.float32), freqs)
    return torch.cos(ang).to(dtype), torch.sin(ang).to(dtype)


def apply_rope(x, cos, sin):
    """apply_rotary_emb, models/base_blocks.py:314-324.  x: [B,T,nh,hd]; interleaved pairs."""
    B, T, nh, hd = x.shape
    xr = x.reshape(B, T, nh, hd // 2, 2)
    x0, x1 = xr[..., 0], xr[..., 1]
    c = cos[:T].view(1, T, 1, hd // 2)
    s = sin[:T].view(1, T, 1, hd // 2)
    return torch.stack([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1).reshape(B, T, nh, hd)


def masked_attention(x_q, x_kv, sd: State, name: str, n_heads: int, src_len: int = 0,
                     attn_dropout_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0):
    """MaskedMultiHeadAttention (x_q is x_kv) / MaskedMultiHeadCrossAttention,
    models/base_blocks.py:175-203 / :267-295.  Causal: key k allowed iff k <= q + src_len."""
    B, T, Cdim = x_q.shape
    hd = Cdim // n_heads
    q = linear(x_q, sd, name + ".q").view(B, T, n_heads, hd)
    k = linear(x_kv, sd, name + ".k").view(B, T, n_heads, hd)
    v = linear(x_kv, sd, name + ".v").view(B, T, n_heads, hd)
    cos, sin = rope_table(hd, T, dtype=x_q.dtype)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * hd ** -0.5
    allowed = torch.tril(torch.ones(T, T, dtype=torch.bool), diagonal=src_len)
    att = att.masked_fill(~allowed, float("-inf"))
    att = torch.softmax(att, dim=-1)
    if attn_dropout_mask is not None:
        att = att * attn_dropout_mask / (1.0 - dropout_p)
    out = (att @ v).transpose(1, 2).reshape(B, T, Cdim)
    return F.linear(out, sd[name + ".projection.weight"])


def mlp(x, sd: State, name: str, drop_mult: Optional[torch.Tensor] = None):
    """MLP.forward with num_layers None/1: Linear → nn.LayerNorm → GELU → Linear → Dropout,
    models/base_blocks.py:22-26, 44-47.  `drop_mult` = the train-mode dropout as an explicit multiplier
    tensor (mask / (1 - p)); None = eval mode / p = 0."""
    h = linear(x, sd, name + ".layers.0")
    h = layer_norm(h, sd[name + ".layers.1.weight"], sd[name + ".layers.1.bias"])
    out = linear(gelu(h), sd, name + ".layers.3")
    return out if drop_mult is None else out * drop_mult


# ----------------------------------------------------------------------------- temporal model
def temporal_block(xs: List[torch.Tensor], ib, sd: State, prefix: str, *, n_heads: int,
                   ln_type: str, src_len: int = 0, taps: Optional[dict] = None,
                   drop: Optional[dict] = None, layer: int = 0):
    """BaseBlockTemporal.forward + SEABlockTemporal exchange with add_info_after_cross=True,
    ib_scale_mode='mlp', ib_addition_mode='add' (the only mode either config selects)."""
    V = len(xs)
    xs = list(xs)
    # train-mode dropout as explicit multipliers (mask / (1-p)) per site: ("self", l, i), ("cross", l, i, j)
    # on the attention probabilities [B,nh,T,T]; ("mlp", l, i), ("tipi", l, i) on [B,T,E]
    dm = (lambda *k: drop.get(k)) if drop is not None else (lambda *k: None)

    def attn(xq, xkv, name, mult):
        if mult is None:
            return masked_attention(xq, xkv, sd, name, n_heads, src_len)
        return masked_attention(xq, xkv, sd, name, n_heads, src_len, attn_dropout_mask=mult, dropout_p=0.0)

    # per-field causal self-attention, models/temporal.py:135-136
    for i in range(V):
        n = norm(xs[i], ib, sd, f"{prefix}.ln.exp.{i}.0", ln_type)
        xs[i] = xs[i] + attn(n, n, f"{prefix}.attn.self.{i}", dm("self", layer, i))
        if taps is not None:
            taps[f"self.{i}"] = xs[i]
    # State-Exchange Attention, sequential over i (Gauss–Seidel), models/temporal.py:176-192
    for i in range(V):
        acc = 0
        for j in range(V):
            if j == i:
                continue
            di = linear(xs[i], sd, f"{prefix}.cross_down.{i}")
            dj = linear(xs[j], sd, f"{prefix}.cross_down.{j}")
            ni = norm(di, ib, sd, f"{prefix}.ln_cross.{i}", ln_type)
            nj = norm(dj, ib, sd, f"{prefix}.ln_cross.{j}", ln_type)
            a = attn(ni, nj, f"{prefix}.cross_attn.{i}.{j}", dm("cross", layer, i, j))
            acc = acc + linear(gelu(a), sd, f"{prefix}.cross_up.{i}")
        xs[i] = xs[i] + acc
        if taps is not None:
            taps[f"exchange.{i}"] = xs[i]
    # TIPI after the exchange, one shared MLP(ib_num → scale·ib_num → E), models/temporal.py:140-142
    for i in range(V):   # the shared ib-MLP is CALLED once per stream: its dropout draws a mask per call
        xs[i] = xs[i] + mlp(ib, sd, f"{prefix}.ib", dm("tipi", layer, i))
    # stream MLP + proj (proj REPLACES the stream), models/temporal.py:144-146
    for i in range(V):
        n = norm(xs[i], ib, sd, f"{prefix}.ln.exp.{i}.2", ln_type)
        xs[i] = xs[i] + mlp(n, sd, f"{prefix}.mlp.{i}", dm("mlp", layer, i))
        xs[i] = linear(xs[i], sd, f"{prefix}.proj.{i}")
        if taps is not None:
            taps[f"block_out.{i}"] = xs[i]
    return xs


def temporal_forward(x, ib, sd: State, *, num_layers: int, n_heads: int, ln_type: str,
                     src_len: int = 0, taps: Optional[dict] = None, drop: Optional[dict] = None):
    """TemporalModel.forward, models/temporal.py:405-416.  x [B,T,V,E], ib [B,T,ib_num]."""
    V = x.shape[2]
    xs = [x[:, :, i, :] for i in range(V)]
    for layer in range(num_layers):
        xs = temporal_block(xs, ib, sd, f"blocks.{layer}", n_heads=n_heads, ln_type=ln_type,
                            src_len=src_len, taps=taps, drop=drop, layer=layer)
    xs = [norm(xs[i], ib, sd, f"ln.{i}", ln_type) for i in range(V)]
    return torch.stack(xs, dim=2)


def rollout(x0, ib, steps: int, sd: State, **kw):
    """Autoregressive loop of utils/train_utils.py:202-209: full prefix recomputed every step."""
    seq = x0
    for i in range(steps):
        out = temporal_forward(seq, ib[:, : i + 1], sd, **kw)
        seq = torch.cat([seq, out[:, -1:]], dim=1)
    return seq[:, 1:]


# ------------------------------------------------------------------------------ spatial model
def positional_encoding(length: int, d_model: int, dtype=torch.float32):
    """PositionalEncoding buffer, models/base_blocks.py:360-369."""
    pe = torch.zeros(length, d_model)
    pos = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div[: d_model // 2])
    return pe.to(dtype)


def encoder_block(z, sd: State, prefix: str, n_heads: int):
    """EncoderBlock / MultiHeadAttention (non-causal, no RoPE), models/base_blocks.py:105-138."""
    B, P, Es = z.shape
    hd = Es // n_heads
    n = layer_norm(z, sd[f"{prefix}.ln_exp1_1.weight"])
    q = linear(n, sd, f"{prefix}.attn_1.q").view(B, P, n_heads, hd).transpose(1, 2)
    k = linear(n, sd, f"{prefix}.attn_1.k").view(B, P, n_heads, hd).transpose(1, 2)
    v = linear(n, sd, f"{prefix}.attn_1.v").view(B, P, n_heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, P, Es)
    z = z + F.linear(o, sd[f"{prefix}.attn_1.projection.weight"])
    n = layer_norm(z, sd[f"{prefix}.ln_exp1_2.weight"])
    return z + mlp(n, sd, f"{prefix}.mlp_1")


def spatial_encode(x, sd: State, *, field_groups: Sequence[Sequence[int]], num_layers: int,
                   n_heads: int, prefix: str = "encode"):
    """PointwiseEncode.forward, models/encoder_decoder.py:105-123.  x [B,P,F,C] → [B,P,G,D]."""
    B, P, _, _ = x.shape
    zs = []
    for g, group in enumerate(field_groups):
        xg = x[:, :, list(group), :].reshape(B, P, 1, -1)  # field-major flattening
        h = gelu(F.linear(xg, sd[f"{prefix}.encoders.{g}.layer1.weight"]))
        zs.append(linear(h, sd, f"{prefix}.encoders.{g}.layer2"))
    z = torch.cat(zs, dim=-2).reshape(B, P, -1)
    z = z + positional_encoding(P, z.shape[-1], z.dtype)
    for layer in range(num_layers):
        z = encoder_block(z, sd, f"{prefix}.blocks.{layer}", n_heads)
    z = layer_norm(z, sd[f"{prefix}.ln.weight"], sd[f"{prefix}.ln.bias"])
    return z.reshape(B, P, len(field_groups), -1)


def spatial_decode(z, sd: State, *, field_groups: Sequence[Sequence[int]], prefix: str = "decode"):
    """Decode.forward, models/encoder_decoder.py:137-146.  z [B,P,G,D] → [B,P,F,C]."""
    B, P, _, _ = z.shape
    outs = []
    for g, group in enumerate(field_groups):
        h = gelu(F.linear(z[:, :, g:g + 1, :], sd[f"{prefix}.decoders.{g}.layer1.weight"]))
        outs.append(linear(h, sd, f"{prefix}.decoders.{g}.layer2").reshape(B, P, len(group), -1))
    return torch.cat(outs, dim=2)


def spatial_forward(x, sd: State, *, field_groups, num_layers: int, n_heads: int, pad_idx=-9999):
    """SpatialModel.forward incl. the in-place pad rewrite, models/encoder_decoder.py:161-176."""
    x[x == pad_idx] = 0.0
    z = spatial_encode(x, sd, field_groups=field_groups, num_layers=num_layers, n_heads=n_heads)
    return spatial_decode(z, sd, field_groups=field_groups)


# ----------------------------------------------------------------- latent layout (next row f-3)
def transform_processed_data(z, tr: int, T: int, P: int, G: int):
    """utils/train_utils.py:315-337: [tr*T, P, G, D] → [tr, T, G, P*D]."""
    D = z.shape[-1]
    return z.reshape(tr, T, P, G, D).permute(0, 1, 3, 2, 4).reshape(tr, T, G, -1)


def inverse_transform_processed_data(y, tr: int, T: int, P: int, G: int):
    """utils/train_utils.py:339-362: [tr, T, G, P*D] → [tr*T, P, G, D]."""
    D = y.shape[-1] // P
    return y.reshape(tr, T, G, P, D).permute(0, 1, 3, 2, 4).reshape(tr * T, P, G, D)


# ------------------------------------------------------------------- synthetic parameter sets
def _normal(shape, gen, std=0.02):
    return torch.randn(shape, generator=gen) * std


def init_temporal_state(*, embed_dim: int, n_heads: int, scale_ratio: int, num_variables: int,
                        down_proj: int = 2, num_layers: int = 1, ln_type: str = "adaln",
                        ib_num: int = 1, seed: int = 42, bias_std: float = 0.0,
                        norm_jitter: float = 0.0) -> State:
    """A state dict with the reference's LIVE parameter names/shapes and its init distribution
    (nn.Linear ~ N(0, 0.02), zero bias, norm weight 1 / bias 0 — models/temporal.py:395-402).
    Dead parameters and the tril / freqs_cis buffers are not materialised.
    ``bias_std`` / ``norm_jitter`` optionally perturb biases and norm affine params so tests
    exercise every term.  The RNG stream is NOT the reference's; goldens carry real reference
    weights where bit-for-bit reference inputs matter."""
    g = torch.Generator().manual_seed(seed)
    E, V = embed_dim, num_variables
    Dd, H = E // down_proj, int(E * scale_ratio)
    sd: State = {}

    def lin(name, n_out, n_in, bias=True):
        sd[name + ".weight"] = _normal((n_out, n_in), g)
        if bias:
            sd[name + ".bias"] = _normal((n_out,), g, bias_std) if bias_std else torch.zeros(n_out)

    def nrm(name, dim, kind, with_bias):
        sd[name + ".weight"] = torch.ones(dim) + (_normal((dim,), g, norm_jitter) if norm_jitter else 0)
        if kind == "adaln":
            sd[name + ".bias"] = _normal((dim,), g, norm_jitter) if norm_jitter else torch.zeros(dim)
            lin(name + ".cond_mlp.0", 2 * dim, ib_num)
            lin(name + ".cond_mlp.2", 2 * dim, 2 * dim)
        elif with_bias:
            sd[name + ".bias"] = _normal((dim,), g, norm_jitter) if norm_jitter else torch.zeros(dim)

    for layer in range(num_layers):
        p = f"blocks.{layer}"
        for i in range(V):
            nrm(f"{p}.ln.exp.{i}.0", E, ln_type, False)
            nrm(f"{p}.ln.exp.{i}.2", E, ln_type, False)
            for nm in ("q", "k", "v"):
                lin(f"{p}.attn.self.{i}.{nm}", E, E)
            lin(f"{p}.attn.self.{i}.projection", E, E, bias=False)
            lin(f"{p}.cross_down.{i}", Dd, E)
            lin(f"{p}.cross_up.{i}", E, Dd)
            nrm(f"{p}.ln_cross.{i}", Dd, ln_type, False)
            for j in range(V):
                if j == i:
                    continue
                for nm in ("q", "k", "v"):
                    lin(f"{p}.cross_attn.{i}.{j}.{nm}", Dd, Dd)
                lin(f"{p}.cross_attn.{i}.{j}.projection", Dd, Dd, bias=False)
            lin(f"{p}.mlp.{i}.layers.0", H, E)
            nrm(f"{p}.mlp.{i}.layers.1", H, "ln", True)
            lin(f"{p}.mlp.{i}.layers.3", E, H)
            lin(f"{p}.proj.{i}", E, E)
        hid = max(1, int(ib_num * scale_ratio))
        lin(f"{p}.ib.layers.0", hid, ib_num)
        nrm(f"{p}.ib.layers.1", hid, "ln", True)
        lin(f"{p}.ib.layers.3", E, hid)
    for i in range(V):
        nrm(f"ln.{i}", E, ln_type, False)
    return sd


def init_spatial_state(*, field_groups, n_inp: int, mlp_hidden: int, num_layers: int,
                       embed_dim: int, seed: int = 42) -> State:
    """SpatialModel parameter names/shapes (variational=False).  Encoder blocks use the
    reference's N(0,0.02) init; the patch MLPs / decoder use a uniform(-1/sqrt(fan_in)) stand-in
    for torch's default Linear init (models/encoder_decoder.py:89-103)."""
    g = torch.Generator().manual_seed(seed)
    G = len(field_groups)
    Es = G * embed_dim
    sd: State = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    for gi, group in enumerate(field_groups):
        cin = n_inp * len(group)
        sd[f"encode.encoders.{gi}.layer1.weight"] = uni((mlp_hidden, cin), cin)
        sd[f"encode.encoders.{gi}.layer2.weight"] = uni((embed_dim, mlp_hidden), mlp_hidden)
        sd[f"encode.encoders.{gi}.layer2.bias"] = uni((embed_dim,), mlp_hidden)
        sd[f"decode.decoders.{gi}.layer1.weight"] = uni((mlp_hidden, embed_dim), embed_dim)
        sd[f"decode.decoders.{gi}.layer2.weight"] = uni((cin, mlp_hidden), mlp_hidden)
        sd[f"decode.decoders.{gi}.layer2.bias"] = uni((cin,), mlp_hidden)
    for layer in range(num_layers):
        p = f"encode.blocks.{layer}"
        sd[f"{p}.ln_exp1_1.weight"] = torch.ones(Es)
        sd[f"{p}.ln_exp1_2.weight"] = torch.ones(Es)
        for nm in ("q", "k", "v"):
            sd[f"{p}.attn_1.{nm}.weight"] = _normal((Es, Es), g)
            sd[f"{p}.attn_1.{nm}.bias"] = torch.zeros(Es)
        sd[f"{p}.attn_1.projection.weight"] = _normal((Es, Es), g)
        sd[f"{p}.mlp_1.layers.0.weight"] = _normal((4 * Es, Es), g)
        sd[f"{p}.mlp_1.layers.0.bias"] = torch.zeros(4 * Es)
        sd[f"{p}.mlp_1.layers.1.weight"] = torch.ones(4 * Es)
        sd[f"{p}.mlp_1.layers.1.bias"] = torch.zeros(4 * Es)
        sd[f"{p}.mlp_1.layers.3.weight"] = _normal((Es, 4 * Es), g)
        sd[f"{p}.mlp_1.layers.3.bias"] = torch.zeros(Es)
    sd["encode.ln.weight"] = torch.ones(Es)
    sd["encode.ln.bias"] = torch.zeros(Es)
    return sd


# The two configs of the reference (configs/cylinder_flow.py:112-128, configs/multiphase_flow.py)
TEMPORAL_CONFIGS = {
    "cylinder_flow": dict(embed_dim=1024, n_heads=8, scale_ratio=8, num_variables=2, down_proj=2,
                          num_layers=1, ln_type="adaln", batch_size=2, src_len=399),
    "multiphase_flow": dict(embed_dim=2048, n_heads=8, scale_ratio=8, num_variables=2, down_proj=2,
                            num_layers=1, ln_type="ln", batch_size=4, src_len=199),
}
SPATIAL_CONFIGS = {
    "cylinder_flow": dict(field_groups=[[0, 1], [2]], mlp_hidden=480, num_layers=12, embed_dim=16,
                          n_heads=8),
    "multiphase_flow": dict(field_groups=[[0, 1], [2]], mlp_hidden=624, num_layers=12, embed_dim=32,
                            n_heads=8),
}

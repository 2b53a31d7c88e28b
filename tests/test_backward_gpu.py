"""Backward parity: per-kernel against torch autograd on the same data, whole model against the
oracle's gradients (goldens from the unmodified reference), and a 100-step training-loss curve."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import golden_recipe as gr
from oracle import sea_oracle as so
from tests.helpers import SEED, rel_l2, temporal_case

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("d", [512, 1024, 2048])
@pytest.mark.parametrize("kind", [0, 1])
def test_norm_bwd(cuda, d, kind):
    from sea_b200 import ops
    M = 301
    g = torch.Generator(device="cuda").manual_seed(d + kind)
    x = (torch.randn(M, d, device=cuda, generator=g) * 1.5 + 0.3).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(d, device=cuda, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(d, device=cuda, generator=g)).requires_grad_(True)
    cond = (torch.randn(M, 2 * d, device=cuda, generator=g) * 0.3).requires_grad_(True)
    dy = torch.randn(M, d, device=cuda, generator=g)
    dres = torch.randn(M, d, device=cuda, generator=g)
    if kind == 1:
        mu, var = x.mean(-1, keepdim=True), x.var(-1, keepdim=True, unbiased=False)
        y = (x - mu) / (var + 1e-5).sqrt() * (w + cond[:, :d] + 1) + (b + cond[:, d:])
    else:
        y = F.layer_norm(x, (d,), w, None, 1e-5)
    (y * dy).sum().backward()
    _, _, st = ops.norm_fwd(x.detach(), w.detach(), bias=b.detach() if kind else None,
                            cond=cond.detach() if kind else None, kind=kind, out_dtype=torch.float32, stats=True)
    dx, dxb, dw, db, dc = ops.norm_bwd(dy, x.detach(), st, w.detach(), cond=cond.detach() if kind else None,
                                       kind=kind, dres=dres, want_bf16=True)
    assert _rel(dx, x.grad + dres) < 1e-5
    assert _rel(dxb.float(), x.grad + dres) < 5e-3
    assert _rel(dw, w.grad) < 1e-4
    if kind:
        assert _rel(db, b.grad) < 1e-4 and _rel(dc, cond.grad) < 1e-5


@pytest.mark.parametrize("H", [256, 8192, 16384])
def test_ln_gelu_bwd(cuda, H):
    from sea_b200 import ops
    M = 77
    g = torch.Generator(device="cuda").manual_seed(H)
    h = (torch.randn(M, H, device=cuda, generator=g) * 1.5).bfloat16()
    w = (1 + 0.1 * torch.randn(H, device=cuda, generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(H, device=cuda, generator=g)).requires_grad_(True)
    dg = torch.randn(M, H, device=cuda, generator=g).bfloat16()
    hf = h.float().requires_grad_(True)
    y = F.gelu(F.layer_norm(hf, (H,), w, b, 1e-5))
    (y * dg.float()).sum().backward()
    _, st = ops.ln_gelu_fwd_with_stats(h, w.detach(), b.detach())
    dh, dw, db = ops.ln_gelu_bwd(dg, h, st, w.detach(), b.detach())
    assert _rel(dh.float(), hf.grad) < 6e-3
    assert _rel(dw, w.grad) < 1e-3 and _rel(db, b.grad) < 1e-3


@pytest.mark.parametrize("H", [256, 8192, 16384])
def test_ln_gelu_bwd_fp32(cuda, H):
    """fp32-parity variant (fp32 dg / h / dh, erff / expf) against autograd in float64."""
    from sea_b200 import ops
    M = 77
    g = torch.Generator(device="cuda").manual_seed(H + 1)
    h = torch.randn(M, H, device=cuda, generator=g) * 1.5
    w = (1 + 0.1 * torch.randn(H, device=cuda, generator=g))
    b = (0.1 * torch.randn(H, device=cuda, generator=g))
    dg = torch.randn(M, H, device=cuda, generator=g)
    hd_, wd, bd = (t.double().requires_grad_(True) for t in (h, w, b))
    (F.gelu(F.layer_norm(hd_, (H,), wd, bd, 1e-5)) * dg.double()).sum().backward()
    _, st = ops.ln_gelu_fwd_with_stats(h, w, b)
    dh, dw, db = ops.ln_gelu_bwd(dg, h, st, w, b)
    assert dh.dtype == torch.float32
    assert _rel(dh, hd_.grad) < 2e-5
    assert _rel(dw, wd.grad) < 2e-5 and _rel(db, bd.grad) < 2e-5


@pytest.mark.parametrize("hd", [32, 64, 128, 256])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,src_len", [(2, 77, 0), (1, 200, 0), (2, 33, 2)])
def test_attention_bwd(cuda, hd, dt, B, T, src_len):
    from sea_b200 import lib, ops
    nh = 2
    g = torch.Generator(device="cuda").manual_seed(hd + T)
    qkv = torch.randn(B * T, 3 * nh * hd, device=cuda, generator=g).to(dt)
    q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:2 * nh * hd], qkv[:, 2 * nh * hd:]
    d_o = torch.randn(B * T, nh * hd, device=cuda, generator=g).to(dt)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(B, T, nh, hd).transpose(1, 2)
    kh = kf.view(B, T, nh, hd).transpose(1, 2)
    vh = vf.view(B, T, nh, hd).transpose(1, 2)
    att = (qh @ kh.transpose(-2, -1)) * hd ** -0.5
    att = att.masked_fill(torch.ones(T, T, device=cuda).tril(diagonal=src_len) == 0, float("-inf"))
    o_ref = (torch.softmax(att, -1) @ vh).transpose(1, 2).reshape(B * T, nh * hd)
    (o_ref * d_o.float()).sum().backward()
    lib.sea_attention_force_simt(1)
    try:
        o, lse = ops.attention_fwd(q, k, v, nh, src_len=src_len, B=B, want_lse=True)
    finally:
        lib.sea_attention_force_simt(0)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, d_o, lse, nh, B=B, src_len=src_len)
    tol = 2e-5 if dt == torch.float32 else 1.5e-2
    assert _rel(dq.float(), qf.grad) < tol
    assert _rel(dk.float(), kf.grad) < tol
    assert _rel(dv.float(), vf.grad) < tol


def _rope_tables(hd, T):
    """freqs (models/base_blocks.py:300-306) as complex [T, hd/2] and as the pair-major real table."""
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
    ang = torch.outer(torch.arange(T, dtype=torch.float32), inv)
    fc = torch.polar(torch.ones_like(ang), ang)
    return fc, torch.view_as_real(fc).float().transpose(0, 1).contiguous()


@pytest.mark.parametrize("hd", [64, 128, 256])
@pytest.mark.parametrize("B,T,src_len,rope", [(2, 128, 0, False), (1, 399, 0, True), (3, 199, 0, True),
                                             (1, 521, 3, False), (2, 1030, 0, True)])
def test_attention_bwd_tensor_core(cuda, hd, B, T, src_len, rope):
    """tcgen05 backward (multi-tile, ragged tails, both roles, RoPE undone in the epilogue) against
    torch autograd through the same attention incl. the rotation (models/base_blocks.py:184-197)."""
    from sea_b200 import ops
    nh = 2
    g = torch.Generator(device="cuda").manual_seed(hd + T)
    pre = (torch.randn(B * T, 3 * nh * hd, device=cuda, generator=g) * 0.8).bfloat16()
    d_o = torch.randn(B * T, nh * hd, device=cuda, generator=g).bfloat16()
    fc, tab = _rope_tables(hd, T)
    fc, tab = fc.to(cuda), tab.to(cuda)
    xf = pre.float().detach().clone().requires_grad_(True)
    q0, k0, v0 = (xf[:, i * nh * hd:(i + 1) * nh * hd].view(B, T, nh, hd) for i in range(3))

    def rot(x):
        xc = torch.view_as_complex(x.float().reshape(B, T, nh, hd // 2, 2))
        return torch.view_as_real(xc * fc[None, :, None, :]).flatten(3)

    qr, kr = (rot(q0), rot(k0)) if rope else (q0, k0)
    # the kernels see bf16 rotated q/k (what the projection GEMM epilogue stores)
    qkv_dev = torch.cat([qr.detach().reshape(B * T, -1), kr.detach().reshape(B * T, -1),
                         v0.detach().reshape(B * T, -1)], dim=1).bfloat16()
    qd, kd, vd = (qkv_dev[:, i * nh * hd:(i + 1) * nh * hd] for i in range(3))
    qh, kh, vh = (t.transpose(1, 2) for t in (qr, kr, v0))
    att = (qh @ kh.transpose(-2, -1)) * hd ** -0.5
    att = att.masked_fill(torch.ones(T, T, device=cuda).tril(diagonal=src_len) == 0, float("-inf"))
    o_ref = (torch.softmax(att, -1) @ vh).transpose(1, 2).reshape(B * T, nh * hd)
    (o_ref * d_o.float()).sum().backward()
    gq, gk, gv = (xf.grad[:, i * nh * hd:(i + 1) * nh * hd] for i in range(3))
    o, lse = ops.attention_fwd(qd, kd, vd, nh, src_len=src_len, B=B, want_lse=True)
    dq, dk, dv = ops.attention_bwd(qd, kd, vd, o, d_o, lse, nh, B=B, src_len=src_len,
                                   rope_table=tab if rope else None)
    torch.cuda.synchronize()
    errs = (_rel(dq.float(), gq), _rel(dk.float(), gk), _rel(dv.float(), gv))
    print(f"\n[attn bwd tc] hd={hd} B={B} T={T} src_len={src_len} rope={rope}: dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e}")
    assert max(errs) < 1.5e-2


def _mirror(tag, ln, cuda):
    from sea_b200.temporal import TemporalModel
    g, sd, cfg, x, ib, tgt, _ = temporal_case(tag, ln, requires_grad=True)
    E, nh, scale, V, B, T, _ = [int(v) for v in g["meta"]]
    m = TemporalModel(1, E, nh, 64, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    m.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=False)
    return g, sd, cfg, m.to(cuda).train(), x, ib, tgt


@pytest.mark.parametrize("tag,ln", [("small_adaln", "adaln"), ("small_ln", "ln"), ("small_v3", "ln")])
def test_model_gradients_match_reference(cuda, tag, ln):
    g, sd, cfg, m, x, ib, tgt = _mirror(tag, ln, cuda)
    # oracle gradients (fp32 autograd on CPU; pinned to the reference by tests/test_oracle_golden.py)
    x = x.detach().clone().requires_grad_(True)
    loss_ref = F.mse_loss(so.temporal_forward(x, ib, sd, **cfg), tgt)
    loss_ref.backward()
    xg = x.detach().to(cuda).requires_grad_(True)
    y = m(xg, ib.to(cuda))
    loss = F.mse_loss(y, tgt.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    params = dict(m.named_parameters())
    worst = 0.0
    for name in [str(n) for n in g["grad_names"]]:
        got, ref = params[name].grad, sd[name].grad
        assert got is not None, name
        ref_norm = ref.double().norm().item()
        if ref_norm < 2e-5:   # e.g. the TIPI input layer: LayerNorm over 2-8 values, pure cancellation
            continue
        err = _rel(got.cpu(), ref)
        cos = F.cosine_similarity(got.cpu().flatten().double(), ref.flatten().double(), dim=0).item()
        worst = max(worst, err)
        assert cos > 0.995 and err < 8e-2, (name, err, cos)
    for name in [str(n) for n in g["dead_params"]]:
        assert params[name].grad is None, name
    assert _rel(xg.grad.cpu(), x.grad) < 5e-2
    print(f"\n[temporal bwd] {tag}: loss {loss.item():.6f} vs {loss_ref.item():.6f}; worst param-grad rel err {worst:.3e}")


@pytest.mark.parametrize("tag,ln", [("cylinder_flow", "adaln"), ("multiphase_flow", "ln")])
def test_full_width_gradients_match_reference_golden(cuda, tag, ln):
    """The two real configs at full width (E = 1024 / 2048: tcgen05 attention backward at head dim 128 / 256 self and
    64 / 128 cross, K = 8192 / 16384 weight-gradient GEMMs) against the gradient goldens the UNMODIFIED reference
    produced (oracle/make_golden.py): loss, the norm of every live parameter gradient, its projection on a seeded
    random probe vector (error of the projection ~ ||g - g_ref||), and every small gradient element-wise."""
    from sea_b200.temporal import TemporalModel
    from tests.helpers import load_golden
    g = load_golden("temporal_" + tag)
    E, nh, scale, V, B, T, _ = [int(v) for v in g["meta"]]
    shapes = gr.temporal_shapes(embed_dim=E, n_heads=nh, scale_ratio=scale, num_variables=V, ln_type=ln)
    sd = gr.fill_state(shapes, SEED)
    x, ib, tgt = gr.temporal_inputs(B, T, V, E, SEED)
    m = TemporalModel(1, E, nh, 64, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    m.load_state_dict(sd, strict=False)
    m = m.to(cuda).train()
    loss = F.mse_loss(m(x.to(cuda), ib.to(cuda)), tgt.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) < 2e-3 * float(g["loss"])
    params = dict(m.named_parameters())
    worst_norm = worst_probe = worst_small = 0.0
    for name, ref_norm, ref_probe in zip(g["grad_names"], g["grad_norms"], g["grad_probes"]):
        name = str(name)
        got = params[name].grad
        assert got is not None, name
        if ref_norm < 1e-7:
            continue
        e_norm = abs(got.double().norm().item() - ref_norm) / ref_norm
        pv = gr.probe_vector(name, got.shape, SEED).to(cuda).double()
        e_probe = abs((got.double() * pv).sum().item() - ref_probe) / ref_norm
        worst_norm, worst_probe = max(worst_norm, e_norm), max(worst_probe, e_probe)
        assert e_norm < 3e-2 and e_probe < 8e-2, (name, e_norm, e_probe)
        if "grad:" + name in g.files:
            e = _rel(got.cpu(), torch.from_numpy(g["grad:" + name]))
            worst_small = max(worst_small, e)
            assert e < 5e-2, (name, e)
    for name in [str(n) for n in g["dead_params"]]:
        assert params[name].grad is None, name
    print(f"\n[full-width bwd golden] {tag}: loss {loss.item():.6f} vs {float(g['loss']):.6f}; worst |norm| err "
          f"{worst_norm:.2e}, worst probe err {worst_probe:.2e}, worst small-gradient rel err {worst_small:.2e}")


def test_gradient_accumulation_semantics(cuda):
    g, sd, cfg, m, x, ib, tgt = _mirror("small_ln", "ln", cuda)
    xc, ic, tc = x.detach().to(cuda), ib.to(cuda), tgt.to(cuda)
    F.mse_loss(m(xc, ic), tc).backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    F.mse_loss(m(xc, ic), tc).backward()          # no zero_grad: must accumulate
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert _rel(p.grad, 2 * g1[n]) < 1e-3, n
    m.zero_grad(set_to_none=True)
    F.mse_loss(m(xc, ic), tc).backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert _rel(p.grad, g1[n]) < 1e-3, n


@pytest.mark.parametrize("tag,ln", [("small_adaln", "adaln"), ("small_ln", "ln")])
def test_training_loss_curve_100_steps(cuda, tag, ln):
    """train/train_temporal.py:252-260 semantics (AdamW lr 1e-4, MSE, dropout 0): per-step loss of the
    CUDA path vs the fp32 oracle trained on CPU from the same init and data."""
    g, sd, cfg, m, x, ib, tgt = _mirror(tag, ln, cuda)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    opt_ref = torch.optim.AdamW(list(leaves.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    xd, ibd, td = x.detach(), ib, tgt
    xc, ic, tc = xd.to(cuda), ibd.to(cuda), td.to(cuda)
    ref_losses, losses = [], []
    for step in range(100):
        opt_ref.zero_grad()
        lr_ = F.mse_loss(so.temporal_forward(xd, ibd, leaves, **cfg), td)
        lr_.backward()
        opt_ref.step()
        ref_losses.append(lr_.item())
        opt.zero_grad()
        lo = F.mse_loss(m(xc, ic), tc)
        lo.backward()
        opt.step()
        losses.append(lo.item())
    ref_losses, losses = np.array(ref_losses), np.array(losses)
    rel = np.abs(losses - ref_losses) / np.abs(ref_losses)
    print(f"\n[train 100 steps] {tag}: loss {ref_losses[0]:.4f} -> {ref_losses[-1]:.4f} (oracle), "
          f"{losses[0]:.4f} -> {losses[-1]:.4f} (cuda); max per-step rel diff {rel.max():.3e}")
    assert ref_losses[-1] < 0.9 * ref_losses[0]          # it actually trains
    assert rel.max() < 2e-2


def test_gradient_accumulation_and_fresh_overwrite(cuda):
    """zero_grad(set_to_none) -> backward takes the 'fresh' path (weight-gradient GEMMs overwrite, only
    the small-gradient region is zero-filled); a second backward without zero_grad must accumulate
    (2x); a later fresh backward must not see stale values."""
    g, sd, cfg, m, x, ib, tgt = _mirror("small_adaln", "adaln", cuda)
    xg, ibg, tg = x.to(cuda), ib.to(cuda), tgt.to(cuda)

    def run():
        F.mse_loss(m(xg, ibg), tg).backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    m.engine().flat_grad().fill_(123.0)          # poison: the fresh path must not read it
    for p in m.parameters():
        p.grad = None
    g1 = run()
    g2 = run()                                    # accumulates on top of g1
    for p in m.parameters():
        p.grad = None
    g3 = run()
    assert len(g1) > 50
    for n in g1:
        assert torch.isfinite(g1[n]).all(), n
        scale = g1[n].abs().max().item() + 1e-12
        assert (g2[n] - 2 * g1[n]).abs().max().item() <= 2e-3 * scale + 1e-7, n      # atomics reorder sums
        assert (g3[n] - g1[n]).abs().max().item() <= 2e-3 * scale + 1e-7, n


@pytest.mark.parametrize("V,L,ln", [(4, 2, "adaln"), (3, 2, "ln")])
def test_more_streams_and_layers_match_oracle(cuda, V, L, ln):
    """BASELINE configs[4] sweeps the stream count (V = 2, 3, 4); the reference also allows several blocks.
    Forward, loss and every live parameter gradient of a V-stream, L-layer model against the fp32 oracle on
    fresh seeded inputs (tensor-core path: head dim 64, cross head dim 32 falls to the CUDA-core attention)."""
    from sea_b200.temporal import TemporalModel
    E, nh, scale, B, T = 128, 2, 2, 2, 21
    shapes = gr.temporal_shapes(embed_dim=E, n_heads=nh, scale_ratio=scale, num_variables=V, num_layers=L, ln_type=ln)
    sd = gr.fill_state(shapes, 23)
    for v in sd.values():
        v.requires_grad_(True)
    x, ib, tgt = gr.temporal_inputs(B, T, V, E, 23)
    m = TemporalModel(L, E, nh, 64, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    missing = m.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=False)
    assert not missing.unexpected_keys
    m = m.to(cuda).train()
    xr = x.clone().requires_grad_(True)
    loss_ref = F.mse_loss(so.temporal_forward(xr, ib, sd, num_layers=L, n_heads=nh, ln_type=ln), tgt)
    loss_ref.backward()
    xg = x.to(cuda).requires_grad_(True)
    y = m(xg, ib.to(cuda))
    loss = F.mse_loss(y, tgt.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    params = dict(m.named_parameters())
    checked = 0
    for name, ref_p in sd.items():
        ref = ref_p.grad
        if ref is None or ref.double().norm().item() < 2e-5:
            continue
        got = params[name].grad
        assert got is not None, name
        cos = F.cosine_similarity(got.cpu().flatten().double(), ref.flatten().double(), dim=0).item()
        assert cos > 0.99 and _rel(got.cpu(), ref) < 0.12, (name, _rel(got.cpu(), ref), cos)
        checked += 1
    assert checked > 40 * L
    assert _rel(xg.grad.cpu(), xr.grad) < 6e-2

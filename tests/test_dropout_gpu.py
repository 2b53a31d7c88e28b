"""Train-mode nn.Dropout inside the kernels (models/base_blocks.py:47, 194, 286).  The masks are
counter-based functions of (seed, site, element); sea_dropout_mask exports exactly what the fused kernels
apply, so the oracle can be run with the identical masks: forward, loss and every gradient must agree
as tightly as without dropout."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import sea_oracle as so
from tests.helpers import temporal_case

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _mask(seed, site, n, p, dev):
    from sea_b200 import check, lib
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.sea_dropout_mask(C.c_uint64(seed), C.c_uint32(site), C.c_int64(n), C.c_float(p),
                               C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mask")
    return out


def _site(l, kind, i, j):
    return ((l * 4 + kind) * 4 + i) * 4 + j


def test_mask_statistics(cuda):
    for p in (0.1, 0.5):
        m = _mask(1234, 7, 1 << 22, p, cuda).float()
        assert abs(m.mean().item() - (1 - p)) < 2e-3
        # neighbours / different sites / different seeds are uncorrelated
        assert abs(((m[1:] - (1 - p)) * (m[:-1] - (1 - p))).mean().item()) < 2e-3
        m2 = _mask(1234, 8, 1 << 22, p, cuda).float()
        assert abs(((m - (1 - p)) * (m2 - (1 - p))).mean().item()) < 2e-3
        assert not torch.equal(m, _mask(1235, 7, 1 << 22, p, cuda).float())


@pytest.mark.parametrize("hd,B,T", [(64, 2, 77), (128, 1, 399), (256, 2, 130), (128, 2, 128)])
def test_attention_dropout_fwd_bwd(cuda, hd, B, T):
    """probability dropout in the tcgen05 forward (one- and two-tile kernels) and backward against torch
    autograd with the exported mask."""
    from sea_b200 import ops
    from sea_b200 import _structs as S
    from sea_b200._lib import check, lib
    nh, p, seed, site = 2, 0.1, 99991 + T, 5
    g = torch.Generator(device="cuda").manual_seed(hd + T)
    qkv = (torch.randn(B * T, 3 * nh * hd, device=cuda, generator=g) * 0.8).bfloat16()
    q, k, v = (qkv[:, i * nh * hd:(i + 1) * nh * hd] for i in range(3))
    d_o = torch.randn(B * T, nh * hd, device=cuda, generator=g).bfloat16()
    T2 = (T + 1) & ~1
    mult = _mask(seed, site, B * nh * T * T2, p, cuda).view(B, nh, T, T2)[..., :T].float() / (1 - p)
    xf = qkv.float().detach().clone().requires_grad_(True)
    qh, kh, vh = (xf[:, i * nh * hd:(i + 1) * nh * hd].view(B, T, nh, hd).transpose(1, 2) for i in range(3))
    att = (qh @ kh.transpose(-2, -1)) * hd ** -0.5
    att = att.masked_fill(torch.ones(T, T, device=cuda).tril() == 0, float("-inf"))
    o_ref = ((torch.softmax(att, -1) * mult) @ vh).transpose(1, 2).reshape(B * T, nh * hd)
    (o_ref * d_o.float()).sum().backward()
    gq, gk, gv = (xf.grad[:, i * nh * hd:(i + 1) * nh * hd] for i in range(3))

    o = torch.empty(B * T, nh * hd, device=cuda, dtype=torch.bfloat16)
    lse = torch.empty(B, nh, T, device=cuda)
    a = S.AttnArgs()
    a.q, a.k, a.v, a.ldq, a.ldk, a.ldv = q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), k.stride(0), v.stride(0)
    a.o, a.ldo, a.lse, a.B, a.T, a.n_heads, a.head_dim, a.src_len = o.data_ptr(), o.stride(0), lse.data_ptr(), B, T, nh, hd, 0
    a.scale, a.prec, a.dropout_p, a.dropout_site, a.dropout_seed = hd ** -0.5, 0, p, site, seed
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.sea_attention_fwd(C.byref(a), st), "attention_fwd")
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, nh, T, device=cuda)
    bw = S.AttnBwdArgs()
    bw.q, bw.k, bw.v, bw.o, bw.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr()
    bw.ldq, bw.ldk, bw.ldv, bw.ldo, bw.lddo = q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0)
    bw.lse, bw.delta, bw.dq, bw.dk, bw.dv = lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    bw.lddq, bw.lddk, bw.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    bw.B, bw.T, bw.n_heads, bw.head_dim, bw.src_len, bw.scale, bw.prec = B, T, nh, hd, 0, hd ** -0.5, 0
    bw.dropout_p, bw.dropout_site, bw.dropout_seed = p, site, seed
    check(lib.sea_attention_bwd(C.byref(bw), st), "attention_bwd")
    torch.cuda.synchronize()
    errs = (_rel(o.float(), o_ref), _rel(dq.float(), gq), _rel(dk.float(), gk), _rel(dv.float(), gv))
    print(f"\n[attention dropout] hd={hd} B={B} T={T}: o {errs[0]:.2e} dq {errs[1]:.2e} dk {errs[2]:.2e} dv {errs[3]:.2e}")
    assert max(errs) < 1.5e-2


@pytest.mark.parametrize("ln,V", [("adaln", 2), ("ln", 2), ("ln", 3)])
def test_model_with_dropout_matches_oracle_with_same_masks(cuda, ln, V):
    from oracle import golden_recipe as gr
    from sea_b200.temporal import TemporalModel
    # tensor-core head dims (128 self / 64 exchange), like both reference configs; the oracle itself is pinned
    # to the reference by tests/test_oracle_golden.py
    E, nh, scale, B, T = 256, 2, 4, 2, 40
    shapes = gr.temporal_shapes(embed_dim=E, n_heads=nh, scale_ratio=scale, num_variables=V, ln_type=ln)
    sd = {k: v.requires_grad_(True) if v.dtype.is_floating_point else v for k, v in gr.fill_state(shapes, 17).items()}
    x, ib, tgt = gr.temporal_inputs(B, T, V, E, 17)
    cfg = dict(num_layers=1, n_heads=nh, ln_type=ln)
    tag = f"{ln}-V{V}"
    p = 0.1
    m = TemporalModel(1, E, nh, 64, scale, 0, V, 2, p, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    m.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=False)
    m = m.to(cuda).train()
    torch.manual_seed(321)
    y = m(x.detach().to(cuda), ib.to(cuda))
    loss = F.mse_loss(y, tgt.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    seed = m.engine().last_dropout_seed
    assert seed != 0
    # the identical masks for the oracle
    T2, M = (T + 1) & ~1, B * T
    drop = {}
    for i in range(V):
        drop[("self", 0, i)] = (_mask(seed, _site(0, 0, i, 0), B * nh * T * T2, p, cuda).view(B, nh, T, T2)[..., :T].float() / (1 - p)).cpu()
        drop[("mlp", 0, i)] = (_mask(seed, _site(0, 2, i, 0), M * E, p, cuda).view(B, T, E).float() / (1 - p)).cpu()
        drop[("tipi", 0, i)] = (_mask(seed, _site(0, 3, i, 0), M * E, p, cuda).view(B, T, E).float() / (1 - p)).cpu()
        for j in range(V):
            if j != i:
                drop[("cross", 0, i, j)] = (_mask(seed, _site(0, 1, i, j), B * nh * T * T2, p, cuda)
                                            .view(B, nh, T, T2)[..., :T].float() / (1 - p)).cpu()
    xr = x.detach().clone().requires_grad_(True)
    y_ref = so.temporal_forward(xr, ib, sd, drop=drop, **cfg)
    loss_ref = F.mse_loss(y_ref, tgt)
    loss_ref.backward()
    y_eval = so.temporal_forward(x.detach(), ib, {k: v.detach() for k, v in sd.items()}, **cfg)
    assert _rel(y_ref.detach(), y_eval) > 5e-2            # dropout really changes the output ...
    e_y = _rel(y.detach().cpu(), y_ref.detach())
    assert e_y < 2e-2, e_y                                 # ... and the kernels follow the same masks
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    params = dict(m.named_parameters())
    worst = 0.0
    for name, got_p in params.items():
        got, ref = got_p.grad, (sd[name].grad if name in sd else None)
        if ref is None:
            assert got is None, name       # the reference's dead parameters stay without gradient
            continue
        assert got is not None, name
        if ref.double().norm().item() < 2e-5:
            continue
        err = _rel(got.cpu(), ref)
        cos = F.cosine_similarity(got.cpu().flatten().double(), ref.flatten().double(), dim=0).item()
        if "ib.layers.0." in name or "ib.layers.1." in name:
            # TIPI input layer: LayerNorm over 8 values of a scalar — its gradient is a difference of nearly
            # equal terms (the no-dropout parity test skips it for the same reason); direction must agree
            assert cos > 0.99 and err < 0.25, (name, err, cos)
            continue
        worst = max(worst, err)
        assert cos > 0.995 and err < 8e-2, (name, err, cos)
    print(f"\n[model dropout] {tag}: forward rel {e_y:.2e}, loss {loss.item():.5f} vs {loss_ref.item():.5f}, worst grad rel {worst:.2e}")
    # eval mode: no dropout, repeatable
    m.eval()
    with torch.no_grad():
        y1, y2 = m(x.detach().to(cuda), ib.to(cuda)), m(x.detach().to(cuda), ib.to(cuda))
    assert torch.equal(y1, y2) and _rel(y1.cpu(), y_eval) < 2e-2

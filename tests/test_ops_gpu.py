"""Per-kernel parity of the HBM-bound kernels and attention against torch fp32 on the same data."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("d", [64, 512, 1024, 2048])
@pytest.mark.parametrize("kind", [0, 1])
def test_norm_fwd(cuda, d, kind):
    from sea_b200 import ops
    M = 203
    g = torch.Generator(device="cuda").manual_seed(d + kind)
    big = torch.randn(M, 2, d, device=cuda, generator=g) * 2 + 0.5
    x = big[:, 1]                                   # strided stream view
    w = 1 + 0.1 * torch.randn(d, device=cuda, generator=g)
    b = 0.1 * torch.randn(d, device=cuda, generator=g)
    cond = torch.randn(M, 2 * d, device=cuda, generator=g) * 0.3
    if kind == 1:
        mu, var = x.mean(-1, keepdim=True), x.var(-1, keepdim=True, unbiased=False)
        ref = (x - mu) / (var + 1e-5).sqrt() * (w + cond[:, :d] + 1) + (b + cond[:, d:])
        y, _, st = ops.norm_fwd(x, w, bias=b, cond=cond, kind=1, out_dtype=torch.float32, stats=True)
    else:
        ref = F.layer_norm(x, (d,), w, None, 1e-5)
        y, _, st = ops.norm_fwd(x, w, kind=0, out_dtype=torch.float32, stats=True)
    assert _rel(y, ref) < 2e-6
    assert _rel(st[:, 0], x.mean(-1)) < 1e-5
    yb, _, _ = ops.norm_fwd(x, w, bias=b if kind else None, cond=cond if kind else None, kind=kind)
    assert _rel(yb.float(), ref) < 4e-3


def test_norm_fwd_with_tipi(cuda):
    from sea_b200 import ops
    M, d = 97, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(M, d, device=cuda, generator=g)
    ib = torch.rand(M, 1, device=cuda, generator=g)
    w0, b0 = torch.randn(8, 1, device=cuda, generator=g), torch.randn(8, device=cuda, generator=g) * 0.1
    lw, lb = 1 + 0.1 * torch.randn(8, device=cuda, generator=g), 0.1 * torch.randn(8, device=cuda, generator=g)
    w3, b3 = torch.randn(d, 8, device=cuda, generator=g) * 0.2, torch.randn(d, device=cuda, generator=g) * 0.1
    w = 1 + 0.1 * torch.randn(d, device=cuda, generator=g)
    gh = ops.tipi_hidden(ib, w0, b0, lw, lb)
    ref_g = F.gelu(F.layer_norm(F.linear(ib, w0, b0), (8,), lw, lb, 1e-5))
    assert _rel(gh, ref_g) < 1e-5
    y, x_out, _ = ops.norm_fwd(x, w, kind=0, out_dtype=torch.float32, tipi=(gh, w3, b3))
    x_ref = x + F.linear(ref_g, w3, b3)
    assert _rel(x_out, x_ref) < 1e-6
    assert _rel(y, F.layer_norm(x_ref, (d,), w, None, 1e-5)) < 2e-6


def test_adaln_hidden(cuda):
    from sea_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(6)
    ib = torch.rand(50, 1, device=cuda, generator=g)
    w1, b1 = torch.randn(256, 1, device=cuda, generator=g), torch.randn(256, device=cuda, generator=g)
    ref = F.silu(F.linear(ib, w1, b1))
    assert _rel(ops.adaln_hidden(ib, w1, b1, torch.float32), ref) < 1e-6
    assert _rel(ops.adaln_hidden(ib, w1, b1).float(), ref) < 4e-3


@pytest.mark.parametrize("H", [256, 8192, 16384])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_ln_gelu(cuda, H, dt):
    from sea_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(H)
    h = (torch.randn(33, H, device=cuda, generator=g) * 1.5 + 0.2).to(dt)
    w = 1 + 0.1 * torch.randn(H, device=cuda, generator=g)
    b = 0.1 * torch.randn(H, device=cuda, generator=g)
    out = ops.ln_gelu_fwd(h, w, b)
    ref = F.gelu(F.layer_norm(h.float(), (H,), w, b, 1e-5))
    assert _rel(out.float(), ref) < (2e-6 if dt == torch.float32 else 4e-3)


def test_pack_and_split_gemm_is_fp32_accurate(cuda):
    """3-way bf16 split along K makes the bf16 tensor-core GEMM match an fp64 product to ~1e-6."""
    from sea_b200 import ops
    M, N, K = 300, 256, 512
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(M, K, device=cuda, generator=g)
    b = torch.randn(N, K, device=cuda, generator=g) * 0.02
    a6, b6 = ops.pack_operand(a, split=1), ops.pack_operand(b, split=2)
    out = torch.empty(M, N, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(a6, b6, out_f32=out)], M, N, 6 * K)
    ref = (a.double() @ b.double().t())
    # tensor-core fp32 accumulation is not round-to-nearest: ~6e-6 measured, bar is 1e-4
    assert _rel(out, ref) < 2e-5
    # transposed + split weights reproduce dY @ W
    bt6 = ops.pack_operand(b, transpose=True, split=2)          # [K, 6N]
    dy = torch.randn(M, N, device=cuda, generator=g)
    dy6 = ops.pack_operand(dy, split=1)
    dx = torch.empty(M, K, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(dy6, bt6, out_f32=dx)], M, K, 6 * N)
    assert _rel(dx, dy.double() @ b.double()) < 2e-5
    # plain cast / transpose
    assert torch.equal(ops.pack_operand(a), a.bfloat16())
    assert torch.equal(ops.pack_operand(a, transpose=True), a.t().contiguous().bfloat16())
    assert _rel(ops.pack_operand(a, act=ops.ACT_GELU).float(), F.gelu(a)) < 4e-3


def _attn_ref(q, k, v, B, nh, src_len=0):
    M, Cd = q.shape
    T, hd = M // B, Cd // nh
    qh = q.float().view(B, T, nh, hd).transpose(1, 2)
    kh = k.float().view(B, T, nh, hd).transpose(1, 2)
    vh = v.float().view(B, T, nh, hd).transpose(1, 2)
    att = (qh @ kh.transpose(-2, -1)) * hd ** -0.5
    mask = torch.ones(T, T, device=q.device).tril(diagonal=src_len) == 0
    att = att.masked_fill(mask, float("-inf"))
    lse = torch.logsumexp(att, dim=-1)
    o = (torch.softmax(att, -1) @ vh).transpose(1, 2).reshape(M, Cd)
    return o, lse


@pytest.mark.parametrize("hd", [32, 64, 128, 256])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,src_len", [(2, 77, 0), (1, 399, 0), (3, 16, 2), (2, 1, 0)])
def test_attention_simt(cuda, hd, dt, B, T, src_len):
    from sea_b200 import lib, ops
    nh = 2
    g = torch.Generator(device="cuda").manual_seed(hd + T)
    qkv = torch.randn(B * T, 3 * nh * hd, device=cuda, generator=g).to(dt)
    q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:2 * nh * hd], qkv[:, 2 * nh * hd:]
    lib.sea_attention_force_simt(1)
    try:
        o, lse = ops.attention_fwd(q, k, v, nh, src_len=src_len, B=B, want_lse=True)
    finally:
        lib.sea_attention_force_simt(0)
    ref, ref_lse = _attn_ref(q, k, v, B, nh, src_len)
    assert _rel(o.float(), ref) < (3e-6 if dt == torch.float32 else 6e-3)
    assert _rel(lse, ref_lse) < 1e-5


@pytest.mark.parametrize("hd", [64, 128, 256])
@pytest.mark.parametrize("B,T,src_len", [(2, 77, 0), (1, 399, 0), (3, 128, 0), (2, 1, 0), (1, 640, 0), (2, 200, 3)])
def test_attention_tcgen05(cuda, hd, B, T, src_len):
    from sea_b200 import ops
    nh = 2
    g = torch.Generator(device="cuda").manual_seed(hd + T)
    qkv = torch.randn(B * T, 3 * nh * hd, device=cuda, generator=g).bfloat16()
    q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:2 * nh * hd], qkv[:, 2 * nh * hd:]
    from sea_b200._lib import lib
    lib.sea_attention_small(0)          # T <= 128 would otherwise go to the short-sequence kernel
    try:
        o, lse = ops.attention_fwd(q, k, v, nh, src_len=src_len, B=B, want_lse=True)
    finally:
        lib.sea_attention_small(1)
    torch.cuda.synchronize()
    ref, ref_lse = _attn_ref(q, k, v, B, nh, src_len)
    assert torch.isfinite(o.float()).all()
    assert _rel(o.float(), ref) < 8e-3
    assert _rel(lse, ref_lse) < 1e-4


@pytest.mark.parametrize("hd", [64, 128])
@pytest.mark.parametrize("T", [1, 2, 5, 16, 17, 33, 64, 100, 127, 128])
@pytest.mark.parametrize("src_len", [0, 2])
def test_attention_short_sequences(cuda, hd, T, src_len):
    """attention_small.cu (T <= 128: the rollout's prefixes) against the fp32 reference and the tcgen05 kernel."""
    from sea_b200 import ops
    from sea_b200._lib import lib
    B, nh = 3, 4
    g = torch.Generator(device="cuda").manual_seed(1000 * hd + 10 * T + src_len)
    qkv = torch.randn(B * T, 3 * nh * hd + 8, device=cuda, generator=g).bfloat16()       # padded pitch
    q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:2 * nh * hd], qkv[:, 2 * nh * hd:3 * nh * hd]
    try:
        lib.sea_attention_small(2)      # every T <= 128 (the default hands T > 24 / 40 to the tcgen05 kernel)
        o, lse = ops.attention_fwd(q, k, v, nh, src_len=src_len, B=B, want_lse=True)
        lib.sea_attention_small(0)
        o_tc, lse_tc = ops.attention_fwd(q, k, v, nh, src_len=src_len, B=B, want_lse=True)
    finally:
        lib.sea_attention_small(1)
    torch.cuda.synchronize()
    ref, ref_lse = _attn_ref(q, k, v, B, nh, src_len)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert _rel(o.float(), ref) < 8e-3
    assert _rel(lse, ref_lse) < 1e-4
    assert _rel(o.float(), o_tc.float()) < 8e-3
    assert _rel(lse, lse_tc) < 1e-4


def test_attention_tcgen05_peaked_scores(cuda):
    """Large score range exercises the lazy-rescale path (running max grows by > 2^8)."""
    from sea_b200 import ops
    B, T, nh, hd = 1, 512, 2, 128
    g = torch.Generator(device="cuda").manual_seed(11)
    q = (torch.randn(B * T, nh * hd, device=cuda, generator=g) * 3).bfloat16()
    k = (torch.randn(B * T, nh * hd, device=cuda, generator=g) * 3).bfloat16()
    k[300:] *= 2.0
    v = torch.randn(B * T, nh * hd, device=cuda, generator=g).bfloat16()
    o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
    ref, ref_lse = _attn_ref(q, k, v, B, nh)
    assert _rel(o.float(), ref) < 1e-2
    assert _rel(lse, ref_lse) < 1e-4


@pytest.mark.parametrize("k_chunk", [0, 512])
def test_split_gemm_long_k(cuda, k_chunk):
    """K = 8192 (the MLP down-projection): chunked accumulation keeps the split product at fp32 accuracy."""
    from sea_b200 import ops
    M, N, K = 256, 256, 8192
    g = torch.Generator(device="cuda").manual_seed(10)
    a = torch.randn(M, K, device=cuda, generator=g).abs()      # same-sign terms: worst case for drift
    b = torch.randn(N, K, device=cuda, generator=g).abs() * 0.02
    bias = torch.randn(N, device=cuda, generator=g)
    res = torch.randn(M, N, device=cuda, generator=g)
    a6, b6 = ops.pack_operand(a, split=1), ops.pack_operand(b, split=2)
    out = torch.empty(M, N, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(a6, b6, bias=bias, residual=res, out_f32=out)], M, N, 6 * K, k_chunk=k_chunk)
    ref = a.double() @ b.double().t() + bias.double() + res.double()
    err = _rel(out, ref)
    sgemm = _rel(a @ b.t() + bias + res, ref)
    print(f"\n[split gemm K=8192] k_chunk={k_chunk}: rel err {err:.2e} (torch fp32 matmul: {sgemm:.2e})")
    assert err < (3e-6 if k_chunk else 1e-4)

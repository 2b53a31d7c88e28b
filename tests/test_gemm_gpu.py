"""K1 parity: tcgen05 GEMM + fused epilogues vs torch (fp32 math on bf16-rounded operands)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b):
    return a.float() @ b.float().t()


@pytest.mark.parametrize("bn", [64, 128, 192, 256, 0])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (798, 1024, 1024), (796, 3072, 512),
                                   (37, 72, 136), (1000, 8192, 1024)])
def test_gemm_plain(cuda, bn, M, N, K):
    from sea_b200 import lib, ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=cuda)
    lib.sea_gemm_force_tile_n(bn)
    try:
        ops.gemm_bf16_tn([ops.gemm_problem(a, b, out_f32=out)], M, N, K)
    finally:
        lib.sea_gemm_force_tile_n(0)
    torch.cuda.synchronize()
    ref = _ref(a, b)
    err = (out - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-4, err


def test_gemm_epilogue_bias_residual_gelu(cuda):
    from sea_b200 import ops
    M, N, K = 798, 1024, 512
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    res = torch.randn(M, N, device=cuda, generator=g)
    out = torch.empty(M, N, device=cuda)
    pre = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    act = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    ops.gemm_bf16_tn([ops.gemm_problem(a, b, bias=bias, residual=res, act=ops.ACT_GELU,
                                       out_f32=out, out_pre_bf16=pre, out_bf16=act)], M, N, K)
    torch.cuda.synchronize()
    ref = _ref(a, b) + bias + res
    assert (out - ref).abs().max().item() < 5e-3
    assert (pre.float() - ref).abs().max().item() < 5e-2
    assert (act.float() - torch.nn.functional.gelu(ref)).abs().max().item() < 5e-2


def test_gemm_grouped_and_strided(cuda):
    from sea_b200 import ops
    M, N, K = 400, 512, 256
    g = torch.Generator(device="cuda").manual_seed(2)
    big = torch.randn(M, 2, K, device=cuda, generator=g).bfloat16()   # two interleaved streams
    w = [(torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16() for _ in range(2)]
    out = torch.empty(M, 2, N, device=cuda)
    probs = [ops.gemm_problem(big[:, i], w[i], out_f32=out[:, i]) for i in range(2)]
    ops.gemm_bf16_tn(probs, M, N, K)
    torch.cuda.synchronize()
    for i in range(2):
        ref = _ref(big[:, i], w[i])
        assert (out[:, i] - ref).abs().max().item() < 5e-3


def test_gemm_rope_epilogue(cuda):
    from sea_b200 import ops
    B, T, E, hd = 2, 100, 512, 128
    M, N, K = B * T, 3 * E, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    freqs = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=cuda).float() / hd))
    ang = torch.outer(torch.arange(T, device=cuda).float(), freqs)
    table = torch.stack([ang.cos(), ang.sin()], dim=-1).transpose(0, 1).contiguous()    # pair-major [hd/2, T, 2]
    out = torch.empty(M, N, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(a, b, rope_table=table, rope_cols=2 * E, head_dim=hd,
                                       seq_len=T, out_f32=out)], M, N, K)
    torch.cuda.synchronize()
    ref = _ref(a, b)
    qk = ref[:, :2 * E].reshape(B, T, 2 * E // hd, hd // 2, 2)
    c = torch.view_as_complex(qk.contiguous()) * torch.polar(torch.ones_like(ang), ang)[None, :, None, :]
    ref_rot = torch.cat([torch.view_as_real(c).reshape(M, 2 * E), ref[:, 2 * E:]], dim=1)
    assert (out - ref_rot).abs().max().item() < 5e-3


def test_gemm_throughput_report(cuda):
    """Not a pass/fail perf gate — prints TFLOP/s so the first GPU run tells us where we are."""
    from sea_b200 import lib, ops
    for (M, N, K) in [(8192, 8192, 8192), (12768, 8192, 1024), (12768, 1024, 8192), (798, 8192, 1024)]:
        a = torch.randn(M, K, device=cuda).bfloat16()
        b = torch.randn(N, K, device=cuda).bfloat16()
        out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
        for bn in (128, 256):
            lib.sea_gemm_force_tile_n(bn)
            prob = [ops.gemm_problem(a, b, out_bf16=out)]
            for _ in range(3):
                ops.gemm_bf16_tn(prob, M, N, K)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                ops.gemm_bf16_tn(prob, M, N, K)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            print(f"\n[gemm] M={M} N={N} K={K} BN={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
        lib.sea_gemm_force_tile_n(0)
        ref = _ref(a[:64], b[:64])
        assert (out[:64, :64].float() - ref).abs().max().item() < 0.02 * ref.abs().max().item() + 0.5


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1024, 512, 798), (3072, 1024, 796), (520, 8192, 333), (64, 2048, 1600)])
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_gemm_mn_major_operands(cuda, M, N, K, bn):
    """mn_major = 1: C[M,N] = A^T B with A stored [K,M], B stored [K,N] (the wgrad layout
    dW = dY^T X), accumulated onto an fp32 residual, for every tile width."""
    from sea_b200 import lib, ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(K, M, device=cuda, generator=g).bfloat16()
    b = torch.randn(K, N, device=cuda, generator=g).bfloat16()
    acc = torch.randn(M, N, device=cuda, generator=g)
    ref = acc.double() + a.double().t() @ b.double()
    out = acc.clone()
    lib.sea_gemm_force_tile_n(bn)
    try:
        ops.gemm_bf16_tn([ops.gemm_problem(a, b, residual=out, out_f32=out, mn_major=3)], M, N, K)
    finally:
        lib.sea_gemm_force_tile_n(0)
    torch.cuda.synchronize()
    err = ((out.double() - ref).norm() / ref.norm()).item()
    assert err < 2e-6, err


@pytest.mark.parametrize("M,N,K", [(798, 1024, 3072), (796, 2048, 16384), (100, 512, 1024), (3200, 8192, 1024)])
def test_gemm_b_mn_major(cuda, M, N, K):
    """mn_major = SEA_GEMM_B_MN: C[M,N] = A[M,K] B[K,N] with B stored [K,N] — dgrad reads W [N_out,K_in]
    directly (here K plays N_out and N plays K_in)."""
    from sea_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = (torch.randn(K, N, device=cuda, generator=g) * 0.05).bfloat16()
    out = torch.empty(M, N, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(a, b, out_f32=out, mn_major=2, b_is_static=True)], M, N, K)
    torch.cuda.synchronize()
    ref = a.double() @ b.double()
    err = ((out.double() - ref).norm() / ref.norm()).item()
    # K-major reference run of the same product: the tensor core's fp32 accumulator is not round-to-nearest
    # (drift grows with K), so the bar is "same as the K-major path", not fp32 epsilon
    out2 = torch.empty(M, N, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(a, b.t().contiguous(), out_f32=out2)], M, N, K)
    err2 = ((out2.double() - ref).norm() / ref.norm()).item()
    print(f"\n[gemm B MN-major] M={M} N={N} K={K}: rel err {err:.2e} (K-major path {err2:.2e})")
    assert err < max(2e-6, 1.5 * err2), (err, err2)
    assert torch.equal(out, out2)   # same MMA sequence along K -> bit-identical


@pytest.mark.parametrize("M,N,K,groups", [(32, 1024, 8192, 2), (160, 1024, 8192, 2), (800, 1024, 8192, 2),
                                          (3200, 1024, 8192, 2), (796, 2048, 16384, 1), (100, 512, 1024, 1),
                                          (3200, 8192, 1024, 2), (2000, 520, 2048, 1)])
@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_gemm_stream_k(cuda, M, N, K, groups, bn):
    """Stream-K (contiguous (tile, k-block) ranges per CTA, last arriver reduces the partials in CTA
    order) against the data-parallel schedule of the same kernel: fused epilogue (bias + residual +
    GELU) included; deterministic across repeats; counters left clean."""
    from sea_b200 import lib, ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    ws = torch.zeros(65536 + 148 * 2 * 128 * 256 * 4, dtype=torch.uint8, device=cuda)
    outs = {}
    a = [torch.randn(M, K, device=cuda, generator=g).bfloat16() for _ in range(groups)]
    b = [(torch.randn(N, K, device=cuda, generator=g) * 0.02).bfloat16() for _ in range(groups)]
    bias = torch.randn(N, device=cuda, generator=g)
    res = torch.randn(M, N, device=cuda, generator=g)
    for mode in (0, 2, 2):
        of = [torch.empty(M, N, device=cuda) for _ in range(groups)]
        ob = [torch.empty(M, N, device=cuda, dtype=torch.bfloat16) for _ in range(groups)]
        lib.sea_gemm_stream_k(mode)
        lib.sea_gemm_force_tile_n(bn)
        assert lib.sea_gemm_set_workspace(ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel())) == 0
        try:
            ops.gemm_bf16_tn([ops.gemm_problem(a[i], b[i], bias=bias, residual=res, out_f32=of[i], out_bf16=ob[i],
                                               act=ops.ACT_GELU) for i in range(groups)], M, N, K)
        finally:
            lib.sea_gemm_stream_k(1)
            lib.sea_gemm_force_tile_n(0)
            lib.sea_gemm_set_workspace(None, ctypes.c_size_t(0))
        torch.cuda.synchronize()
        outs.setdefault(mode, []).append((of, ob))
    assert int(ws[:65536].view(torch.int32).abs().sum()) == 0          # arrival counters self-cleaned
    (dp,), (sk1, sk2) = outs[0], outs[2]
    ref = a[0].double() @ b[0].double().t() + bias.double() + res.double()
    err = ((sk1[0][0].double() - ref).norm() / ref.norm()).item()
    assert err < 3e-5, err
    for i in range(groups):
        assert torch.equal(sk1[0][i], sk2[0][i]) and torch.equal(sk1[1][i], sk2[1][i])     # deterministic
        # shorter accumulation chains summed in fp32 RN vs one long chain in the (non-RN) tensor-core
        # accumulator: the two schedules differ at the level of that accumulator's drift
        d = ((sk1[0][i].double() - dp[0][i].double()).norm() / dp[0][i].double().norm()).item()
        assert d < 3e-5, d


@pytest.mark.parametrize("M,N,K,groups,mn", [(3200, 8192, 1024, 2, 0), (796, 2048, 2048, 1, 0), (1000, 1024, 512, 2, 2),
                                             (384, 512, 4096, 1, 0), (1024, 768, 520, 1, 3)])
def test_gemm_cluster_multicast_variant(cuda, M, N, K, groups, mn):
    """Two-CTA clusters with TMA-multicast B halves (opt-in) against the single-CTA schedule of the same
    tile width: bit-identical (same MMA sequence per output element), odd M-tile counts included."""
    from sea_b200 import lib, ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a_shape, b_shape = ((K, M) if mn & 1 else (M, K)), ((K, N) if mn & 2 else (N, K))
    a = [torch.randn(*a_shape, device=cuda, generator=g).bfloat16() for _ in range(groups)]
    b = [(torch.randn(*b_shape, device=cuda, generator=g) * 0.05).bfloat16() for _ in range(groups)]
    bias = torch.randn(N, device=cuda, generator=g)
    outs = []
    for cl in (0, 1):
        of = [torch.empty(M, N, device=cuda) for _ in range(groups)]
        lib.sea_gemm_force_tile_n(256)
        lib.sea_gemm_cluster(cl)
        try:
            ops.gemm_bf16_tn([ops.gemm_problem(a[i], b[i], bias=bias, out_f32=of[i], mn_major=mn, b_is_static=True)
                              for i in range(groups)], M, N, K)
        finally:
            lib.sea_gemm_force_tile_n(0)
            lib.sea_gemm_cluster(0)
        torch.cuda.synchronize()
        outs.append(of)
    for i in range(groups):
        assert torch.equal(outs[0][i], outs[1][i])
    A = a[0].double().t() if mn & 1 else a[0].double()
    Bm = b[0].double() if mn & 2 else b[0].double().t()
    ref = A @ Bm + bias.double()
    assert ((outs[1][0].double() - ref).norm() / ref.norm()).item() < 1e-5


def test_gemm_chunked_inplace_accumulate(cuda):
    """fp32-mode weight gradients: out_f32 aliases residual (dW += ...) with the K loop cut into 512-column chunks."""
    from sea_b200 import ops
    M, N, K = 300, 264, 2000
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    acc = torch.randn(M, N, device=cuda, generator=g)
    want = acc.double() + a.double() @ b.double().t()
    ops.gemm_bf16_tn([ops.gemm_problem(a, b, residual=acc, out_f32=acc)], M, N, K, k_chunk=512)
    torch.cuda.synchronize()
    assert (acc.double() - want).abs().max().item() < 2e-4 * want.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(798, 1024, 512), (199, 3072, 1024)])
def test_split_wgrad_is_fp32_accurate(cuda, M, N, K):
    """dW = dy^T a through the 3x-bf16 split (transposed split packs, contraction over the padded row count, fresh
    accumulator every 512 columns): fp32-accurate against a float64 product (the recipe of linear_bwd_fp32)."""
    import ctypes as C
    from sea_b200 import _structs as S, check, lib, ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    dy = torch.randn(M, N, device=cuda, generator=g)
    a = torch.randn(M, K, device=cuda, generator=g)
    Mp = (M + 7) // 8 * 8
    P1 = torch.zeros(N, 6 * Mp, device=cuda, dtype=torch.bfloat16)
    P2 = torch.zeros(K, 6 * Mp, device=cuda, dtype=torch.bfloat16)

    def pack(src, dst, split):
        pa = S.PackArgs()
        pa.src_f32, pa.ld, pa.R, pa.C = src.data_ptr(), src.stride(0), src.shape[0], src.shape[1]
        pa.transpose, pa.split, pa.act, pa.split_inner = 1, split, 0, Mp
        pa.dst, pa.ld_dst = dst.data_ptr(), dst.stride(0)
        check(lib.sea_pack_operand(C.byref(pa), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pack")

    pack(dy, P1, 1)
    pack(a, P2, 2)
    dW = torch.empty(N, K, device=cuda)
    ops.gemm_bf16_tn([ops.gemm_problem(P1, P2, out_f32=dW)], N, K, 6 * Mp, k_chunk=512)
    torch.cuda.synchronize()
    want = dy.double().t() @ a.double()
    rel = ((dW.double() - want).norm() / want.norm()).item()
    print(f"\n[split wgrad] M={M} N={N} K={K}: rel {rel:.2e}")
    assert rel < 5e-6

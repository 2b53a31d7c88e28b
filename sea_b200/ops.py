"""One thin Python wrapper per C entry point of include/sea_b200.h.

torch is used for device memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from ._lib import GemmEpilogue, GemmProblem, check, lib

ACT_NONE, ACT_GELU = 0, 1


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _ld(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else int(t.stride(0))


def gemm_problem(a, b, *, bias=None, residual=None, gelu_grad_of=None, act=ACT_NONE,
                 rope_table=None, rope_cols=0, head_dim=0, seq_len=0, rope_sign=1.0,
                 out_f32=None, out_pre_bf16=None, out_bf16=None, b_is_static=False, mn_major=0,
                 rope_pos0=0, dropout_p=0.0, dropout_site=0, dropout_seed=0) -> GemmProblem:
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(1) == 1 and b.stride(1) == 1
    # rope_table: pair-major [head_dim/2, rope_ld, 2]
    rope_ld = 0 if rope_table is None else int(rope_table.shape[1])
    e = GemmEpilogue(
        _ptr(bias), _ptr(residual), _ld(residual), _ptr(gelu_grad_of), _ld(gelu_grad_of), act,
        rope_cols, head_dim, seq_len, float(rope_sign), rope_ld, _ptr(rope_table),
        _ptr(out_f32), _ld(out_f32), _ptr(out_pre_bf16), _ld(out_pre_bf16),
        _ptr(out_bf16), _ld(out_bf16), int(rope_pos0), 0, 0, float(dropout_p), int(dropout_site), int(dropout_seed))
    return GemmProblem(_ptr(a), a.stride(0), _ptr(b), b.stride(0), e, int(b_is_static), int(mn_major))


def gemm_bf16_tn(problems: Sequence[GemmProblem], M: int, N: int, K: int, k_chunk: int = 0) -> None:
    arr = (GemmProblem * len(problems))(*problems)
    check(lib.sea_gemm_bf16_tn_chunked(len(problems), arr, M, N, K, k_chunk, _stream()), "gemm_bf16_tn")


# ------------------------------------------------------------------ elementwise / attention ops
from . import _structs as S  # noqa: E402


def norm_fwd(x, weight, *, bias=None, cond=None, kind=0, out_dtype=torch.bfloat16,
             tipi=None, stats=False):
    """x [M,d] fp32 (row stride arbitrary) -> (y, x_out or None, stats or None)."""
    M, d = x.shape
    y = torch.empty(M, d, device=x.device, dtype=out_dtype)
    a = S.NormArgs()
    a.x, a.ldx, a.M, a.d, a.kind = x.data_ptr(), x.stride(0), M, d, kind
    a.weight = weight.data_ptr()
    a.bias = None if bias is None else bias.data_ptr()
    if cond is not None:
        a.cond, a.ldc = cond.data_ptr(), cond.stride(0)
    x_out = None
    if tipi is not None:
        g, w3, b3 = tipi
        x_out = torch.empty(M, d, device=x.device, dtype=torch.float32)
        a.tipi_g, a.tipi_hid, a.tipi_w, a.tipi_b = g.data_ptr(), g.shape[1], w3.data_ptr(), b3.data_ptr()
        a.x_out, a.ldxo = x_out.data_ptr(), d
    if out_dtype == torch.float32:
        a.y_f32, a.ldy_f32 = y.data_ptr(), d
    else:
        a.y_bf16, a.ldy_bf16 = y.data_ptr(), d
    st = torch.empty(M, 2, device=x.device) if stats else None
    a.stats = None if st is None else st.data_ptr()
    check(lib.sea_norm_fwd(C.byref(a), _stream()), "norm_fwd")
    return y, x_out, st


def adaln_hidden(ib, w1, b1, out_dtype=torch.bfloat16):
    M, ib_num = ib.shape
    n = w1.shape[0]
    out = torch.empty(M, n, device=ib.device, dtype=out_dtype)
    ob = _ptr(out) if out_dtype == torch.bfloat16 else None
    of = _ptr(out) if out_dtype == torch.float32 else None
    check(lib.sea_adaln_hidden(_ptr(ib), C.c_int64(ib.stride(0)), M, ib_num, _ptr(w1), _ptr(b1), n, ob, of, _stream()), "adaln_hidden")
    return out


def tipi_hidden(ib, w0, b0, ln_w, ln_b):
    M, ib_num = ib.shape
    hid = w0.shape[0]
    g = torch.empty(M, hid, device=ib.device)
    check(lib.sea_tipi_hidden(_ptr(ib), C.c_int64(ib.stride(0)), M, ib_num, _ptr(w0), _ptr(b0), _ptr(ln_w), _ptr(ln_b), hid,
                              _ptr(g), None, None, _stream()), "tipi_hidden")
    return g


def ln_gelu_fwd(h, weight, bias):
    M, H = h.shape
    g = torch.empty_like(h)
    a = S.LnGeluArgs()
    if h.dtype == torch.bfloat16:
        a.h_bf16, a.g_bf16 = h.data_ptr(), g.data_ptr()
    else:
        a.h_f32, a.g_f32 = h.data_ptr(), g.data_ptr()
    a.ldh, a.ldg, a.M, a.H = h.stride(0), g.stride(0), M, H
    a.weight, a.bias = weight.data_ptr(), bias.data_ptr()
    check(lib.sea_ln_gelu_fwd(C.byref(a), _stream()), "ln_gelu_fwd")
    return g


def pack_operand(src, *, transpose=False, split=0, act=ACT_NONE):
    R, Cc = src.shape
    rows, cols = (Cc, R) if transpose else (R, Cc)
    dst = torch.empty(rows, cols * (6 if split else 1), device=src.device, dtype=torch.bfloat16)
    a = S.PackArgs()
    if src.dtype == torch.float32:
        a.src_f32 = src.data_ptr()
    else:
        a.src_bf16 = src.data_ptr()
    a.ld, a.R, a.C = src.stride(0), R, Cc
    a.transpose, a.split, a.act, a.split_inner = int(transpose), split, act, 0
    a.dst, a.ld_dst = dst.data_ptr(), dst.stride(0)
    check(lib.sea_pack_operand(C.byref(a), _stream()), "pack_operand")
    return dst


def attention_fwd(q, k, v, n_heads, *, src_len=0, B=1, want_lse=False):
    """q,k,v: [B*T, n_heads*hd] views (row stride arbitrary), bf16 or fp32."""
    M, Cdim = q.shape
    T = M // B
    hd = Cdim // n_heads
    o = torch.empty(M, Cdim, device=q.device, dtype=q.dtype)
    lse = torch.empty(B, n_heads, T, device=q.device) if want_lse else None
    a = S.AttnArgs()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
    a.o, a.ldo = o.data_ptr(), o.stride(0)
    a.lse = None if lse is None else lse.data_ptr()
    a.B, a.T, a.n_heads, a.head_dim, a.src_len = B, T, n_heads, hd, src_len
    a.scale = hd ** -0.5
    a.prec = 0 if q.dtype == torch.bfloat16 else 1
    check(lib.sea_attention_fwd(C.byref(a), _stream()), "attention_fwd")
    return (o, lse) if want_lse else o


# ------------------------------------------------------------------------------ backward ops
def norm_bwd(dy, x, stats, weight, *, cond=None, kind=0, dres=None, want_bf16=False):
    """Returns (dx, dx_bf16|None, dweight, dbias|None, dcond|None)."""
    M, d = x.shape
    dx = torch.empty(M, d, device=x.device)
    dxb = torch.empty(M, d, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    dw = torch.zeros(d, device=x.device)
    db = torch.zeros(d, device=x.device) if kind == 1 else None
    dc = torch.empty(M, 2 * d, device=x.device) if kind == 1 else None
    a = S.NormBwdArgs()
    a.dy, a.lddy, a.x, a.ldx, a.stats = dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), stats.data_ptr()
    a.M, a.d, a.kind, a.weight = M, d, kind, weight.data_ptr()
    if cond is not None:
        a.cond, a.ldc = cond.data_ptr(), cond.stride(0)
    if dres is not None:
        a.dres, a.lddres = dres.data_ptr(), dres.stride(0)
    a.dx, a.lddx = dx.data_ptr(), d
    if dxb is not None:
        a.dx_bf16, a.lddx_bf16 = dxb.data_ptr(), d
    a.dweight = dw.data_ptr()
    a.dbias = None if db is None else db.data_ptr()
    if dc is not None:
        a.dcond, a.lddcond, a.dcond_accumulate = dc.data_ptr(), 2 * d, 0
    check(lib.sea_norm_bwd(C.byref(a), _stream()), "norm_bwd")
    return dx, dxb, dw, db, dc


def ln_gelu_bwd(dg, h, stats, weight, bias):
    M, H = h.shape
    dh = torch.empty_like(h)
    dw, db = torch.zeros(H, device=h.device), torch.zeros(H, device=h.device)
    a = S.LnGeluBwdArgs()
    a.dg, a.lddg, a.h, a.ldh, a.stats = dg.data_ptr(), dg.stride(0), h.data_ptr(), h.stride(0), stats.data_ptr()
    a.M, a.H, a.weight, a.bias = M, H, weight.data_ptr(), bias.data_ptr()
    a.dh, a.lddh, a.dweight, a.dbias = dh.data_ptr(), H, dw.data_ptr(), db.data_ptr()
    a.prec = 1 if h.dtype == torch.float32 else 0      # fp32-parity variant: dg / h / dh all fp32
    assert dg.dtype == h.dtype
    check(lib.sea_ln_gelu_bwd(C.byref(a), _stream()), "ln_gelu_bwd")
    return dh, dw, db


def ln_gelu_fwd_with_stats(h, weight, bias):
    M, H = h.shape
    g = torch.empty_like(h)
    st = torch.empty(M, 2, device=h.device)
    a = S.LnGeluArgs()
    if h.dtype == torch.float32:
        a.h_f32, a.g_f32 = h.data_ptr(), g.data_ptr()
    else:
        a.h_bf16, a.g_bf16 = h.data_ptr(), g.data_ptr()
    a.ldh, a.ldg, a.M, a.H = h.stride(0), g.stride(0), M, H
    a.weight, a.bias, a.stats = weight.data_ptr(), bias.data_ptr(), st.data_ptr()
    check(lib.sea_ln_gelu_fwd(C.byref(a), _stream()), "ln_gelu_fwd")
    return g, st


def attention_bwd(q, k, v, o, d_o, lse, n_heads, *, B=1, src_len=0, rope_table=None):
    M, Cd = q.shape
    T, hd = M // B, Cd // n_heads
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, n_heads, T, device=q.device)
    a = S.AttnBwdArgs()
    a.q, a.k, a.v, a.o, a.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0)
    a.lse, a.delta = lse.data_ptr(), delta.data_ptr()
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.lddq, a.lddk, a.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    a.B, a.T, a.n_heads, a.head_dim, a.src_len = B, T, n_heads, hd, src_len
    a.scale = hd ** -0.5
    a.prec = 0 if q.dtype == torch.bfloat16 else 1
    a.rope_table = None if rope_table is None else rope_table.data_ptr()
    a.rope_ld = 0 if rope_table is None else int(rope_table.shape[1])
    check(lib.sea_attention_bwd(C.byref(a), _stream()), "attention_bwd")
    return dq, dk, dv

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
                 out_f32=None, out_pre_bf16=None, out_bf16=None) -> GemmProblem:
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(1) == 1 and b.stride(1) == 1
    e = GemmEpilogue(
        _ptr(bias), _ptr(residual), _ld(residual), _ptr(gelu_grad_of), _ld(gelu_grad_of), act,
        rope_cols, head_dim, seq_len, float(rope_sign), _ptr(rope_table),
        _ptr(out_f32), _ld(out_f32), _ptr(out_pre_bf16), _ld(out_pre_bf16),
        _ptr(out_bf16), _ld(out_bf16))
    return GemmProblem(_ptr(a), a.stride(0), _ptr(b), b.stride(0), e)


def gemm_bf16_tn(problems: Sequence[GemmProblem], M: int, N: int, K: int) -> None:
    arr = (GemmProblem * len(problems))(*problems)
    check(lib.sea_gemm_bf16_tn(len(problems), arr, M, N, K, _stream()), "gemm_bf16_tn")

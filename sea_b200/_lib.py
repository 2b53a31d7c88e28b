"""ctypes binding of include/sea_b200.h (the drop-in C-ABI boundary).

Nothing here touches torch types: callers pass ``tensor.data_ptr()`` and
``torch.cuda.current_stream().cuda_stream``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsea_b200.so")


class LibraryMissing(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("ld_residual", C.c_int64),
        ("gelu_grad_of", C.c_void_p),
        ("ld_gelu", C.c_int64),
        ("act", C.c_int32),
        ("rope_cols", C.c_int32),
        ("head_dim", C.c_int32),
        ("seq_len", C.c_int32),
        ("rope_sign", C.c_float),
        ("rope_ld", C.c_int32),
        ("rope_table", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("ld_out_f32", C.c_int64),
        ("out_pre_bf16", C.c_void_p),
        ("ld_out_pre_bf16", C.c_int64),
        ("out_bf16", C.c_void_p),
        ("ld_out_bf16", C.c_int64),
        ("rope_pos0", C.c_int32),
        ("res_rows_per_batch", C.c_int32),
        ("res_batch_stride", C.c_int64),
        ("dropout_p", C.c_float),
        ("dropout_site", C.c_uint32),
        ("dropout_seed", C.c_uint64),
    ]


class GemmProblem(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("lda", C.c_int64),
        ("b", C.c_void_p),
        ("ldb", C.c_int64),
        ("epi", GemmEpilogue),
        ("b_is_static", C.c_int32),
        ("mn_major", C.c_int32),
    ]


class _Lazy:
    """Loads libsea_b200.so on first attribute access; fails loudly when it is absent."""

    def __init__(self):
        self._dll = None

    def _load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise LibraryMissing(
                    f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'`. "
                    "sea_b200 has no CPU fallback."
                )
            dll = C.CDLL(LIB_PATH)
            dll.sea_strerror.restype = C.c_char_p
            dll.sea_strerror.argtypes = [C.c_int]
            for fn in ("sea_temporal_cache_bytes", "sea_temporal_workspace_bytes", "sea_temporal_cond_cache_bytes",
                       "sea_temporal_kv_cache_bytes", "sea_spatial_cache_bytes"):
                if hasattr(dll, fn):
                    getattr(dll, fn).restype = C.c_size_t
            if os.environ.get("SEA_B200_PDL", "1") == "0":   # A/B switch for programmatic dependent launch
                dll.sea_set_pdl(0)
            if os.environ.get("SEA_B200_CLUSTER", "0") == "1":   # opt-in: two-CTA multicast GEMM variant
                dll.sea_gemm_cluster(1)
            if os.environ.get("SEA_B200_STREAMK", "1") == "0":   # A/B switch: stream-K tails
                dll.sea_gemm_stream_k(0)
            self._dll = dll
        return self._dll

    def __getattr__(self, name):
        return getattr(self._load(), name)


lib = _Lazy()


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib.sea_strerror(int(code)).decode()
        raise RuntimeError(f"sea_b200 {what} failed ({code}): {msg}")

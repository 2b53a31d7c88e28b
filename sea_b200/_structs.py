"""ctypes mirrors of the descriptor structs in include/sea_b200.h (temporal executor)."""
from __future__ import annotations

import ctypes as C

MAX_STREAMS = 4


class Param(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p)]


class NormParams(C.Structure):
    _fields_ = [("weight", Param), ("bias", Param), ("c0_w", Param), ("c0_b", Param),
                ("c2_w", Param), ("c2_b", Param)]


class AttnParams(C.Structure):
    _fields_ = [("q_w", Param), ("q_b", Param), ("k_w", Param), ("k_b", Param),
                ("v_w", Param), ("v_b", Param), ("proj_w", Param)]


class StreamParams(C.Structure):
    _fields_ = [("ln0", NormParams), ("ln2", NormParams), ("ln_cross", NormParams),
                ("self_attn", AttnParams), ("cross_attn", AttnParams * MAX_STREAMS),
                ("down_w", Param), ("down_b", Param), ("up_w", Param), ("up_b", Param),
                ("mlp0_w", Param), ("mlp0_b", Param), ("mlp_ln_w", Param), ("mlp_ln_b", Param),
                ("mlp3_w", Param), ("mlp3_b", Param), ("proj_w", Param), ("proj_b", Param)]


class BlockParams(C.Structure):
    _fields_ = [("s", StreamParams * MAX_STREAMS),
                ("ib0_w", Param), ("ib0_b", Param), ("ib_ln_w", Param), ("ib_ln_b", Param),
                ("ib3_w", Param), ("ib3_b", Param)]


BWD_GROUPS = 5


class TemporalDesc(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("num_streams", C.c_int32),
                ("embed_dim", C.c_int32), ("n_heads", C.c_int32), ("hidden_dim", C.c_int32),
                ("down_dim", C.c_int32), ("ib_num", C.c_int32), ("ib_hidden", C.c_int32),
                ("norm_kind", C.c_int32), ("src_len", C.c_int32), ("max_len", C.c_int32),
                ("precision", C.c_int32), ("ib_time_invariant", C.c_int32), ("cond_cache_valid", C.c_int32),
                ("cond_cache", C.c_void_p), ("cond_cache_bytes", C.c_size_t),
                ("blocks", C.POINTER(BlockParams)),
                ("final_ln", NormParams * MAX_STREAMS),
                ("rope_self", C.c_void_p), ("rope_cross", C.c_void_p),
                ("dropout_p", C.c_float), ("reserved2", C.c_uint32), ("dropout_seed", C.c_uint64),
                ("grads_fresh", C.c_int32), ("splitk_slot", C.c_int32),
                ("grad_f32_base", C.c_void_p), ("grad_bf16", C.c_void_p), ("bwd_events", C.c_void_p * BWD_GROUPS),
                ("aux_stream", C.c_void_p), ("fork_events", C.c_void_p * MAX_STREAMS), ("join_event", C.c_void_p)]


class NormArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ldx", C.c_int64), ("M", C.c_int32), ("d", C.c_int32),
                ("kind", C.c_int32), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("cond", C.c_void_p), ("ldc", C.c_int64), ("cond_div", C.c_int32),
                ("add_rows", C.c_void_p), ("ld_add", C.c_int64), ("add_div", C.c_int32),
                ("tipi_g", C.c_void_p),
                ("tipi_hid", C.c_int32), ("tipi_w", C.c_void_p), ("tipi_b", C.c_void_p),
                ("x_out", C.c_void_p), ("ldxo", C.c_int64), ("y_f32", C.c_void_p),
                ("ldy_f32", C.c_int64), ("y_bf16", C.c_void_p), ("ldy_bf16", C.c_int64),
                ("stats", C.c_void_p), ("cond_folded", C.c_int32), ("x_rows_per_batch", C.c_int32),
                ("x_batch_stride", C.c_int64), ("tipi_dropout_p", C.c_float), ("tipi_dropout_site", C.c_uint32),
                ("tipi_dropout_seed", C.c_uint64)]


class LnGeluArgs(C.Structure):
    _fields_ = [("h_bf16", C.c_void_p), ("h_f32", C.c_void_p), ("ldh", C.c_int64),
                ("M", C.c_int32), ("H", C.c_int32), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("g_bf16", C.c_void_p), ("g_f32", C.c_void_p), ("ldg", C.c_int64),
                ("stats", C.c_void_p)]


class PackArgs(C.Structure):
    _fields_ = [("src_f32", C.c_void_p), ("src_bf16", C.c_void_p), ("ld", C.c_int64),
                ("R", C.c_int32), ("C", C.c_int32), ("transpose", C.c_int32), ("split", C.c_int32),
                ("act", C.c_int32), ("split_inner", C.c_int32), ("dst", C.c_void_p),
                ("ld_dst", C.c_int64)]


class AttnArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
                ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
                ("o", C.c_void_p), ("ldo", C.c_int64), ("lse", C.c_void_p),
                ("B", C.c_int32), ("T", C.c_int32), ("n_heads", C.c_int32), ("head_dim", C.c_int32),
                ("src_len", C.c_int32), ("scale", C.c_float), ("prec", C.c_int32),
                ("dropout_p", C.c_float), ("dropout_site", C.c_uint32), ("dropout_seed", C.c_uint64)]


class NormBwdArgs(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("lddy", C.c_int64), ("x", C.c_void_p), ("ldx", C.c_int64),
                ("stats", C.c_void_p), ("M", C.c_int32), ("d", C.c_int32), ("kind", C.c_int32),
                ("weight", C.c_void_p), ("cond", C.c_void_p), ("ldc", C.c_int64),
                ("dres", C.c_void_p), ("lddres", C.c_int64), ("dx", C.c_void_p), ("lddx", C.c_int64),
                ("dx_bf16", C.c_void_p), ("lddx_bf16", C.c_int64), ("dweight", C.c_void_p),
                ("dbias", C.c_void_p), ("dcond", C.c_void_p), ("lddcond", C.c_int64),
                ("dcond_accumulate", C.c_int32), ("dcond_bf16", C.c_void_p), ("dweight2", C.c_void_p),
                ("dbias2", C.c_void_p)]


class LnGeluBwdArgs(C.Structure):
    _fields_ = [("dg", C.c_void_p), ("lddg", C.c_int64), ("h", C.c_void_p), ("ldh", C.c_int64),
                ("stats", C.c_void_p), ("M", C.c_int32), ("H", C.c_int32), ("weight", C.c_void_p),
                ("bias", C.c_void_p), ("dh", C.c_void_p), ("lddh", C.c_int64),
                ("dweight", C.c_void_p), ("dbias", C.c_void_p), ("prec", C.c_int32)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("d_o", C.c_void_p), ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
                ("ldo", C.c_int64), ("lddo", C.c_int64), ("lse", C.c_void_p), ("delta", C.c_void_p),
                ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("lddq", C.c_int64),
                ("lddk", C.c_int64), ("lddv", C.c_int64), ("B", C.c_int32), ("T", C.c_int32),
                ("n_heads", C.c_int32), ("head_dim", C.c_int32), ("src_len", C.c_int32),
                ("scale", C.c_float), ("prec", C.c_int32), ("rope_table", C.c_void_p), ("rope_ld", C.c_int32),
                ("dropout_p", C.c_float), ("dropout_site", C.c_uint32), ("dropout_seed", C.c_uint64)]


class TipiBwdArgs(C.Structure):
    _fields_ = [("dx", C.c_void_p * MAX_STREAMS), ("lddx", C.c_int64), ("n_streams", C.c_int32),
                ("M", C.c_int32), ("E", C.c_int32), ("hid", C.c_int32), ("ib_num", C.c_int32),
                ("g", C.c_void_p), ("u", C.c_void_p), ("stats", C.c_void_p), ("ib", C.c_void_p),
                ("w3", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("dw3", C.c_void_p), ("db3", C.c_void_p), ("dlnw", C.c_void_p), ("dlnb", C.c_void_p),
                ("dw0", C.c_void_p), ("db0", C.c_void_p)]


class SpatialLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "proj_w", "ln2_w",
                                          "mlp0_w", "mlp0_b", "mlp_ln_w", "mlp_ln_b", "mlp3_w", "mlp3_b")]


class SpatialDesc(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_fields", C.c_int32), ("n_inp", C.c_int32),
                ("n_patches", C.c_int32), ("mlp_hidden", C.c_int32), ("embed_dim", C.c_int32),
                ("n_heads", C.c_int32), ("num_layers", C.c_int32),
                ("group_first_field", C.c_int32 * 4), ("group_num_fields", C.c_int32 * 4),
                ("enc_w1", C.c_void_p * 4), ("enc_w2", C.c_void_p * 4), ("enc_b2", C.c_void_p * 4),
                ("dec_w1", C.c_void_p * 4), ("dec_w2", C.c_void_p * 4), ("dec_b2", C.c_void_p * 4),
                ("layers", C.POINTER(SpatialLayer)), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("pe", C.c_void_p)]

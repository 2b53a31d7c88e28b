"""Host-side mirror of the reference temporal model interface, backed by libsea_b200.so.

``TemporalModel`` has the constructor signature, ``forward(x, x_additional_info)`` signature,
parameter / buffer names and shapes of the reference class (models/temporal.py:326-416), so
checkpoints move both ways and ``train/train_temporal.py:get_model`` can construct either.  The
module tree below only HOLDS parameters; all arithmetic happens in the CUDA library through
``TemporalEngine`` (one FFI call per forward / backward).  ``accelerate`` rebinds ``forward`` on an
*unchanged* reference instance to the same engine.

Supported configuration = what both reference configs select (configs/cylinder_flow.py:112-128):
exchange_mode='sea', ib_scale_mode='mlp', ib_addition_mode='add', add_info_after_cross=True,
LN_type in {'adaln','ln'}.  Anything else raises NotImplementedError — there is no fallback path.
Train-mode dropout (cylinder_flow: 0.1) is applied inside the kernels (counter-based masks).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
import math
import types
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _structs as S
from ._lib import check, lib

PREC = {"bf16": 0, "fp32": 1}


# ----------------------------------------------------------------------------------------------
# Parameter containers (names/shapes only; no arithmetic lives here)
# ----------------------------------------------------------------------------------------------
class _Norm(nn.Module):
    """Holds LayerNorm(weight) [base_blocks.py:80-88] or AdaLN [:330-350] parameters."""

    def __init__(self, dim: int, kind: str, ib_num: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        if kind == "adaln":
            self.bias = nn.Parameter(torch.zeros(dim))
            self.cond_mlp = nn.Sequential(nn.Linear(ib_num, 2 * dim), nn.SiLU(), nn.Linear(2 * dim, 2 * dim))
        else:
            self.bias = None


class _Attention(nn.Module):
    """q/k/v (bias) + projection (no bias) + the reference's buffers (base_blocks.py:155-173)."""

    def __init__(self, n_heads: int, dim: int, max_len: int, src_len: int):
        super().__init__()
        self.n_heads, self.head_dim, self.max_len = n_heads, dim // n_heads, max_len
        self.k = nn.Linear(dim, dim)
        self.q = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.projection = nn.Linear(dim, dim, bias=False)
        hd = self.head_dim
        inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
        ang = torch.outer(torch.arange(max_len, dtype=torch.float32), inv)
        self.register_buffer("freqs_cis", torch.polar(torch.ones_like(ang), ang))
        # kept for state_dict compatibility only; kernels use the predicate k <= q + src_len
        self.register_buffer("tril", torch.ones(max_len, max_len).tril(diagonal=src_len)[None, None])


class _MLP(nn.Module):
    """Linear → LayerNorm → GELU → Linear parameter holder (base_blocks.py:9-47, num_layers 1)."""

    def __init__(self, dim_in: int, scale_ratio, dim_out: Optional[int] = None):
        super().__init__()
        dim_out = dim_in if dim_out is None else dim_out
        self.residual_projection = nn.Linear(dim_in, dim_out) if dim_in != dim_out else None
        hid = max(1, int(dim_in * scale_ratio))
        self.layers = nn.ModuleList([nn.Linear(dim_in, hid), nn.LayerNorm(hid), nn.GELU(), nn.Linear(hid, dim_out)])


class _SinusoidBuffer(nn.Module):
    """`pos_encoder.pe` buffer of the reference block (unused in 'sea' mode, base_blocks.py:355-369)."""

    def __init__(self, d_model: int, max_len: int = 5000):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float32)[:, None]
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div[: d_model // 2])
        self.register_buffer("pe", pe[None])


class _Block(nn.Module):
    def __init__(self, n_heads, max_len, embed_dim, src_len, scale_ratio, num_variables, down_proj,
                 ib_num, kind):
        super().__init__()
        E, V, Dd = embed_dim, num_variables, embed_dim // down_proj
        self.n_heads, self.max_len, self.src_len = n_heads, max_len, src_len
        self.ib = _MLP(ib_num, scale_ratio, E)
        self.ln = nn.ModuleDict({
            "exp": nn.ModuleList([nn.ModuleList([_Norm(E, kind, ib_num) for _ in range(3)]) for _ in range(V)]),
            "cross": _Norm(Dd, kind, ib_num)})
        self.attn = nn.ModuleDict({"self": nn.ModuleList([_Attention(n_heads, E, max_len, src_len) for _ in range(V)])})
        self.mlp = nn.ModuleList([_MLP(E, scale_ratio) for _ in range(V)])
        self.pos_encoder = _SinusoidBuffer(Dd)
        self.proj = nn.ModuleList([nn.Linear(E, E) for _ in range(V)])
        self.cross_down = nn.ModuleList([nn.Linear(E, Dd) for _ in range(V)])
        self.cross_up = nn.ModuleList([nn.Linear(Dd, E) for _ in range(V)])
        self.cross_attn = nn.ModuleList([nn.ModuleList([_Attention(n_heads, Dd, max_len, src_len)
                                                        for _ in range(V)]) for _ in range(V)])
        self.ln_cross = nn.ModuleList([_Norm(Dd, kind, ib_num) for _ in range(V)])


def _check_modes(exchange_mode, ib_scale_mode, ib_addition_mode, add_info_after_cross, LN_type):
    if str(exchange_mode).lower() != "sea":
        raise NotImplementedError(f"sea_b200 accelerates exchange_mode='sea' only (got {exchange_mode!r})")
    if str(ib_scale_mode).lower() != "mlp" or str(ib_addition_mode).lower() != "add":
        raise NotImplementedError("sea_b200 supports ib_scale_mode='mlp', ib_addition_mode='add' only")
    if not add_info_after_cross:
        raise NotImplementedError("sea_b200 supports add_info_after_cross=True only")
    if str(LN_type).lower() not in ("adaln", "ln"):
        raise ValueError(f"Invalid LN_type: {LN_type}. Must be one of {{'adaln', 'ln'}}.")


class TemporalModel(nn.Module):
    """Drop-in for models/temporal.py:TemporalModel (same 17 positional arguments)."""

    def __init__(self, num_layers, embed_dim, n_heads, max_len, scale_ratio, src_len, num_variables,
                 down_proj=2, dropout=0.0, exchange_mode="sea", pos_encoding_mode="learnable",
                 ib_scale_mode="fourier", ib_addition_mode="add", ib_mlp_layers=1, ib_num=1,
                 add_info_after_cross=True, LN_type="adaln", precision: str = "bf16"):
        super().__init__()
        _check_modes(exchange_mode, ib_scale_mode, ib_addition_mode, add_info_after_cross, LN_type)
        if pos_encoding_mode not in ("learnable", "fixed"):
            raise ValueError(f"Invalid pos_encoding_mode '{pos_encoding_mode}'.")
        if ib_mlp_layers not in (None, 1):
            raise NotImplementedError("sea_b200 supports ib_mlp_layers=1 only")
        self.num_variables, self.exchange_mode = num_variables, "sea"
        self.pos_encoding_mode, self.ib_scale_mode, self.ib_addition_mode = pos_encoding_mode, "mlp", "add"
        self.ib_num, self.LN_type, self.dropout_p = ib_num, LN_type, float(dropout)
        kind = LN_type.lower()
        self.blocks = nn.ModuleList([_Block(n_heads, max_len, embed_dim, src_len, scale_ratio,
                                            num_variables, down_proj, ib_num, kind)
                                     for _ in range(num_layers)])
        self.ln = nn.ModuleList([_Norm(embed_dim, kind, ib_num) for _ in range(num_variables)])
        self.apply(self._init_weights)  # models/temporal.py:395-402
        self._precision = precision
        self._engine: Optional[TemporalEngine] = None

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm) or (isinstance(m, _Norm) and m.bias is not None):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def engine(self) -> "TemporalEngine":
        if self._engine is None:
            self._engine = TemporalEngine(self, precision=self._precision, dropout=self.dropout_p)
        return self._engine

    def forward(self, x, x_additional_info):
        assert x.shape[2] == self.num_variables, \
            f"Expected {self.num_variables} variables, but got {x.shape[2]}"
        eng = self.engine()
        eng.dropout = self.dropout_p if self.training else 0.0    # nn.Dropout is the identity in eval mode
        return eng(x, x_additional_info)


# ----------------------------------------------------------------------------------------------
# Engine: builds the C descriptor from a module tree with the reference's naming
# ----------------------------------------------------------------------------------------------
def _P(t: Optional[torch.Tensor], g: Optional[torch.Tensor] = None) -> S.Param:
    return S.Param(None if t is None else t.data_ptr(), None if g is None else g.data_ptr())


class _GradLookup:
    """Maps a parameter tensor to its flat-gradient view (training) or nothing (inference)."""

    def __init__(self, module, views):
        self.by_ptr = {}
        if views:
            for n, p in module.named_parameters():
                if n in views:
                    self.by_ptr[p.data_ptr()] = views[n]

    def __call__(self, t):
        if t is None:
            return S.Param(None, None)
        g = self.by_ptr.get(t.data_ptr())
        return S.Param(t.data_ptr(), None if g is None else g.data_ptr())


class TemporalEngine:
    """Owns the packed-weight cache and workspaces for one module instance."""

    def __init__(self, module: nn.Module, precision: str = "bf16", dropout: float = 0.0):
        self.module = module
        self.precision = precision
        self.dropout = float(dropout)
        self._cache = None
        self._cache_key = None
        self._desc = None
        self._keep = None
        self._ws: Dict[tuple, torch.Tensor] = {}
        self._train_ws: Dict[tuple, list] = {}
        self._flat_grad = None
        self._grad_views = {}
        self.last_launches = 0
        self.total_launches = 0
        # caller-asserted: ib[b,t] == ib[b,0] for all t (set by sea_b200.rollout after checking it)
        self.ib_time_invariant = False
        # rollout(): the condition path's outputs are kept across calls while (B, weights, ib) are fixed
        self.cond_reuse = False
        self._cond_valid = False
        self._cond_buf = None
        self._cond_key = None
        # rollout(): weights cannot change inside the no_grad loop -> skip the per-call freshness scan
        self.weights_frozen = False

    # -- hyper-parameters are read off the module tree (works for the mirror and the reference) --
    def _hyper(self):
        m = self.module
        b0 = m.blocks[0]
        E = b0.proj[0].weight.shape[0]
        att = b0.attn["self"][0]
        kind = "adaln" if hasattr(b0.ln_cross[0], "cond_mlp") else "ln"
        return dict(L=len(m.blocks), V=len(b0.proj), E=E, nh=att.n_heads,
                    H=b0.mlp[0].layers[0].weight.shape[0], Dd=b0.cross_down[0].weight.shape[0],
                    ib_num=b0.ib.layers[0].weight.shape[1], ib_hid=b0.ib.layers[0].weight.shape[0],
                    kind=kind, src_len=int(getattr(b0, "src_len", 0)), max_len=int(att.max_len))

    def _live_params(self):
        """(name, parameter) pairs on the path.  The walk over named_parameters() costs ~0.3 ms of pure Python
        for the configs' models and ran several times per train step, so the list is cached: Parameter OBJECTS
        are stable across .to() / load_state_dict() / optimizer steps (their storage and version are re-read
        on every call); call refresh_parameter_list() after replacing a Parameter object on the module."""
        cached = getattr(self, "_live_cache", None)
        if cached is not None:
            return cached
        self._live_cache = self._walk_live_params()
        return self._live_cache

    def refresh_parameter_list(self) -> None:
        self._live_cache = None
        self._desc = None

    def _walk_live_params(self):
        dead = ("ln.cross.", ".residual_projection.")
        out = []
        for name, p in self.module.named_parameters():
            parts = name.split(".")
            if any(d in name for d in dead):
                continue
            if "ln.exp" in name and parts[parts.index("exp") + 2] == "1":
                continue
            if "cross_attn" in name:
                k = parts.index("cross_attn")
                if parts[k + 1] == parts[k + 2]:
                    continue
            out.append((name, p))
        return out

    # ---- flat gradient buffer: param.grad are views into it (also the DP all-reduce bucket) ----
    @staticmethod
    def _is_gemm_weight(name: str, p: torch.Tensor) -> bool:
        """Parameters whose gradient is produced by a weight-gradient GEMM (every nn.Linear weight on the
        path except the tiny TIPI / cond_mlp.0 layers, whose gradients come from atomics)."""
        return p.dim() == 2 and "ib.layers" not in name and "cond_mlp.0." not in name

    def _ensure_flat_grads(self):
        # every live parameter gets a slot, frozen or not: the backward kernels always have a destination (a frozen
        # parameter's slot is scratch that is never bound to .grad), so partially / fully frozen models just work
        live = list(self._live_params())
        # GEMM weight gradients first, in the order in which the backward FINISHES them (desc.bwd_events groups 0..4), so
        # the data-parallel exchange of each bucket can start while the rest of the backward still runs; the small /
        # atomically accumulated gradients (biases, norm, TIPI, cond_mlp.0) last: final only when the backward ends, and
        # after zero_grad only that tail has to be zero-filled (the weight-gradient GEMMs overwrite, desc.grads_fresh)
        live.sort(key=lambda np_: (not self._is_gemm_weight(*np_), self._bwd_group(np_[0])))
        total = sum((p.numel() + 63) // 64 * 64 for _, p in live)
        dev = live[0][1].device
        if getattr(self, "_flat_grad", None) is None or self._flat_grad.numel() != total or self._flat_grad.device != dev:
            self._flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
            self._flat_grad_bf16 = None
            self._grad_views = {}
            off = 0
            self._small_start = total
            self._group_start = [total] * (S.BWD_GROUPS + 1)     # first element of every weight group, + their end
            for n, p in live:
                if self._is_gemm_weight(n, p):
                    g = self._bwd_group(n)
                    self._group_start[g] = min(self._group_start[g], off)
                else:
                    self._small_start = min(self._small_start, off)
                self._grad_views[n] = self._flat_grad[off:off + p.numel()].view_as(p)
                off += (p.numel() + 63) // 64 * 64
            self._group_start[S.BWD_GROUPS] = self._small_start
            for g in range(S.BWD_GROUPS - 1, -1, -1):             # empty groups collapse onto their successor
                self._group_start[g] = min(self._group_start[g], self._group_start[g + 1])
            self._desc = None  # pointers changed
        return live

    @staticmethod
    def _bwd_group(name: str) -> int:
        """Which backward event (sea_temporal_desc.bwd_events) makes this parameter's gradient final."""
        if name.startswith("ln."):
            return 0                                   # final norms ln.{i}
        parts = name.split(".")
        if parts[0] == "blocks" and parts[1] != "0":
            return 1                                   # upper layers are done before layer 0's first event
        if ".mlp." in name or ".proj." in name:
            return 1
        if ".ln.exp." in name:
            return 2 if parts[parts.index("exp") + 2] == "2" else 4
        if ".ib." in name:
            return 2
        if "cross" in name:
            return 3                                   # cross_down / cross_up / cross_attn / ln_cross
        return 4                                       # attn.self

    def grad_buckets(self):
        """[(begin, end)] element ranges of the flat gradient buffer in exchange order: the weight groups 0..4 as the
        backward finishes them, then the small (reduction-produced) gradients, which are final only at the end."""
        self._ensure_flat_grads()
        gs = self._group_start
        return [(gs[g], gs[g + 1]) for g in range(S.BWD_GROUPS)] + [(self._small_start, self._flat_grad.numel())]

    def flat_grad_bf16(self) -> torch.Tensor:
        """bf16 twin of flat_grad() (allocated on first use): the weight-gradient GEMMs mirror their results into it
        (desc.grad_bf16) — the data-parallel exchange then moves half the bytes."""
        self._ensure_flat_grads()
        if self._flat_grad_bf16 is None:
            self._flat_grad_bf16 = torch.zeros(self._flat_grad.numel(), dtype=torch.bfloat16, device=self._flat_grad.device)
        return self._flat_grad_bf16

    def twin_to_flat_grad(self) -> None:
        """param.grad (fp32 views of flat_grad()) <- the averaged bf16 weight-gradient bucket: what an optimizer other than
        the fused AdamW needs after a bf16 gradient exchange."""
        n_w = self.grad_buckets()[-1][0]
        if n_w:
            with torch.cuda.device(self._flat_grad.device):
                check(lib.sea_cast_bf16_f32(C.c_void_p(self.flat_grad_bf16().data_ptr()), C.c_void_p(self._flat_grad.data_ptr()),
                                            C.c_int64(n_w), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "cast_bf16_f32")

    @staticmethod
    def _is_mlp_weight(name: str) -> bool:
        """blocks.{l}.mlp.{i}.layers.{0,3}.weight (models/temporal.py:80, base_blocks.py:22-26)."""
        return ".mlp." in name and name.endswith(("layers.0.weight", "layers.3.weight"))

    def flat_grad(self) -> torch.Tensor:
        self._ensure_flat_grads()
        return self._flat_grad

    # data-parallel hooks, set by sea_b200.parallel.TrainStep around loss.backward()
    bwd_events = None          # list of BWD_GROUPS cudaEvent_t handles (or None)
    mirror_bf16 = False        # weight-gradient GEMMs also write the bf16 twin

    def anchor_param(self):
        for _, p in self._live_params():
            if p.requires_grad:
                return p
        return None   # every parameter frozen: the backward only produces dL/dx

    def _bind_grads(self):
        """torch semantics: grad None -> zeros; else keep accumulating.  All-None (the state after
        optimizer.zero_grad()) costs a single memset."""
        live = [(n, p) for n, p in self._ensure_flat_grads() if p.requires_grad]
        self._grads_fresh = False
        if all(p.grad is None for _, p in live):
            # zero_grad(set_to_none=True): only the atomically accumulated gradients need zeros; the
            # weight-gradient GEMMs overwrite their destinations on first touch (desc.grads_fresh)
            self._flat_grad[self._small_start:].zero_()
            self._grads_fresh = True
            for n, p in live:
                p.grad = self._grad_views[n]
            return
        for n, p in live:
            v = self._grad_views[n]
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def _build(self, training: bool):
        h = self._hyper()
        m = self.module
        dev = next(m.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("sea_b200 has no CPU path: move the model to a CUDA device")
        for name, p in self._live_params():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"parameter {name} must be contiguous fp32")
        if training:
            self._ensure_flat_grads()
        _P = _GradLookup(m, getattr(self, "_grad_views", None) if training else None)  # noqa: N806
        V, L = h["V"], h["L"]
        if V > S.MAX_STREAMS:
            raise NotImplementedError(f"at most {S.MAX_STREAMS} field streams")
        blocks = (S.BlockParams * L)()

        def norm(mod) -> S.NormParams:
            n = S.NormParams()
            n.weight = _P(mod.weight)
            if hasattr(mod, "cond_mlp"):
                n.bias = _P(mod.bias)
                n.c0_w, n.c0_b = _P(mod.cond_mlp[0].weight), _P(mod.cond_mlp[0].bias)
                n.c2_w, n.c2_b = _P(mod.cond_mlp[2].weight), _P(mod.cond_mlp[2].bias)
            return n

        def attn(mod) -> S.AttnParams:
            a = S.AttnParams()
            a.q_w, a.q_b = _P(mod.q.weight), _P(mod.q.bias)
            a.k_w, a.k_b = _P(mod.k.weight), _P(mod.k.bias)
            a.v_w, a.v_b = _P(mod.v.weight), _P(mod.v.bias)
            a.proj_w = _P(mod.projection.weight)
            return a

        for l, blk in enumerate(m.blocks):
            bp = blocks[l]
            for i in range(V):
                sp = bp.s[i]
                sp.ln0, sp.ln2 = norm(blk.ln["exp"][i][0]), norm(blk.ln["exp"][i][2])
                sp.ln_cross = norm(blk.ln_cross[i])
                sp.self_attn = attn(blk.attn["self"][i])
                for j in range(V):
                    if j != i:
                        sp.cross_attn[j] = attn(blk.cross_attn[i][j])
                sp.down_w, sp.down_b = _P(blk.cross_down[i].weight), _P(blk.cross_down[i].bias)
                sp.up_w, sp.up_b = _P(blk.cross_up[i].weight), _P(blk.cross_up[i].bias)
                ml = blk.mlp[i].layers
                sp.mlp0_w, sp.mlp0_b = _P(ml[0].weight), _P(ml[0].bias)
                sp.mlp_ln_w, sp.mlp_ln_b = _P(ml[1].weight), _P(ml[1].bias)
                sp.mlp3_w, sp.mlp3_b = _P(ml[3].weight), _P(ml[3].bias)
                sp.proj_w, sp.proj_b = _P(blk.proj[i].weight), _P(blk.proj[i].bias)
            il = blk.ib.layers
            bp.ib0_w, bp.ib0_b = _P(il[0].weight), _P(il[0].bias)
            bp.ib_ln_w, bp.ib_ln_b = _P(il[1].weight), _P(il[1].bias)
            bp.ib3_w, bp.ib3_b = _P(il[3].weight), _P(il[3].bias)

        rope_self, rope_cross = self._rope_tables()
        d = S.TemporalDesc()
        d.num_layers, d.num_streams = L, V
        d.embed_dim, d.n_heads, d.hidden_dim, d.down_dim = h["E"], h["nh"], h["H"], h["Dd"]
        d.ib_num, d.ib_hidden = h["ib_num"], h["ib_hid"]
        d.norm_kind = 1 if h["kind"] == "adaln" else 0
        d.src_len, d.max_len = h["src_len"], h["max_len"]
        d.precision = PREC[self.precision]
        d.blocks = C.cast(blocks, C.POINTER(S.BlockParams))
        for i in range(V):
            d.final_ln[i] = norm(m.ln[i])
        d.rope_self, d.rope_cross = rope_self.data_ptr(), rope_cross.data_ptr()
        # two-stream inference schedule (sea_temporal_desc.aux_stream): one auxiliary stream + events per engine
        if self.two_streams and self.precision == "bf16" and V >= 2:
            aux = self._aux_handles(dev, V)
            d.aux_stream = aux[0].cuda_stream
            for i in range(V - 1):
                d.fork_events[i] = aux[1][i]
            d.join_event = aux[1][V - 1]
        self._desc, self._keep, self._h = d, (blocks, rope_self, rope_cross), h
        self._dev = dev

    # Measured on B200 (cylinder_flow, 32 trajectories x 100 steps): 35.4 ms with the fork against 33.8 ms without — the
    # per-stream MLP launches lose more than the overlap with the exchange gains — so the schedule is opt-in.
    two_streams = os.environ.get("SEA_TWO_STREAMS", "0") == "1"

    def _aux_handles(self, dev, V):
        key = (dev.index, V)
        if getattr(self, "_aux", None) is None or self._aux[2] != key:
            evs = []
            with torch.cuda.device(dev):
                for _ in range(V):
                    ev = C.c_void_p()
                    check(lib.sea_event_create(C.byref(ev)), "event_create")
                    evs.append(ev)
                self._aux = (torch.cuda.Stream(device=dev), evs, key)
        return self._aux

    def _rope_tables(self):
        """RoPE tables in the kernels' layout, built ONCE per (engine, source buffers): CUDA graphs recorded by
        RolloutPlan / CachedRolloutPlan / the graphed train step hold their device pointers, so they must
        outlive every _build (train <-> eval switches rebuild the descriptor, not these tensors)."""
        b0 = self.module.blocks[0]
        j0 = 1 if len(b0.proj) > 1 else 0
        fs, fc = b0.attn["self"][0].freqs_cis, b0.cross_attn[0][j0].freqs_cis
        key = (fs.data_ptr(), fc.data_ptr(), fs._version, fc._version, str(fs.device))
        if getattr(self, "_rope_key", None) != key:
            # pair-major [hd/2, max_len, 2]: consecutive positions are contiguous (coalesced epilogue reads)
            self._rope = (torch.view_as_real(fs).float().transpose(0, 1).contiguous(),
                          torch.view_as_real(fc).float().transpose(0, 1).contiguous())
            self._rope_key = key
            self.build_generation = getattr(self, "build_generation", 0) + 1
        return self._rope

    def _key(self, training):
        live = self._live_params()
        return (training, tuple(p.data_ptr() for _, p in live), sum(p._version for _, p in live))

    def _ensure(self, training: bool):
        if self.weights_frozen and self._cache_key is not None and self._cache_key[0] == training:
            return  # rollout(): same weights for every step of the loop, skip the per-call version scan
        key = self._key(training)
        if self._desc is None or self._cache_key is None or key[:2] != self._cache_key[:2]:
            self._build(training)
            nbytes = lib.sea_temporal_cache_bytes(C.byref(self._desc), int(training))
            if self._cache is None or self._cache.numel() < nbytes:
                self._cache = torch.empty(nbytes, dtype=torch.uint8, device=self._dev)
            self._cache_key = None
        if key != self._cache_key:
            with torch.cuda.device(self._dev):
                check(lib.sea_temporal_refresh(C.byref(self._desc), C.c_void_p(self._cache.data_ptr()),
                                               C.c_size_t(self._cache.numel()), int(training),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "temporal_refresh")
            self._cache_key = key

    def after_optimizer_step(self, straight_copies_fresh: bool) -> None:
        """Called by sea_b200.optim.AdamW.step(): the masters changed through raw pointers (no
        ``_version`` bump).  If the optimizer already wrote the bf16 copies, only the fused bias
        vectors (q|k|v, k|v) are re-gathered; otherwise everything is re-packed.  (dgrad reads the same
        [N,K] copy as an MN-major operand, so there are no transposed copies to rebuild.)"""
        if self._cache is None or self._cache_key is None or self._desc is None:
            return
        training = bool(self._cache_key[0])
        what = 4 if straight_copies_fresh else 7   # SEA_REFRESH_BIASES, or SEA_REFRESH_ALL
        with torch.cuda.device(self._dev):
            check(lib.sea_temporal_refresh_ex(C.byref(self._desc), C.c_void_p(self._cache.data_ptr()),
                                              C.c_size_t(self._cache.numel()), int(training), what,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "temporal_refresh_ex")
        self._cache_key = self._key(training)
        self._cond_valid = False

    def workspace(self, B: int, T: int, training: bool) -> torch.Tensor:
        k = (B, T, training)
        ws = self._ws.get(k)
        if ws is None:
            n = lib.sea_temporal_workspace_bytes(C.byref(self._desc), B, T, int(training))
            if len(self._ws) > 8:
                self._ws.clear()
            ws = torch.empty(n, dtype=torch.uint8, device=self._dev)
            self._ws[k] = ws
        return ws

    def acquire_training_workspace(self, B: int, T: int) -> torch.Tensor:
        self._ensure(True)
        pool = self._train_ws.setdefault((B, T), [])
        if pool:
            return pool.pop()
        n = lib.sea_temporal_workspace_bytes(C.byref(self._desc), B, T, 1)
        return torch.empty(n, dtype=torch.uint8, device=self._dev)

    def release_training_workspace(self, key, ws):
        pool = self._train_ws.setdefault(tuple(key), [])
        if len(pool) < 2:
            pool.append(ws)

    @torch.no_grad()
    def backward(self, x, ib, dy, ws, need_dx: bool, dropout_seed: int = 0, dropout_p: float = 0.0):
        if self.precision != "bf16" and dropout_p > 0.0:
            raise NotImplementedError("sea_b200: train-mode dropout runs in bf16 mode only")
        self._ensure(True)
        self._bind_grads()
        B, T, V, E = x.shape
        x = x.contiguous().float()
        ib = ib.contiguous().float()
        dy = dy.contiguous().float()
        dx = torch.empty_like(x) if need_dx else None
        self._desc.grads_fresh = int(getattr(self, "_grads_fresh", False))
        self._desc.dropout_p, self._desc.dropout_seed = float(dropout_p), int(dropout_seed)
        self._desc.grad_f32_base = self._flat_grad.data_ptr()
        self._desc.grad_bf16 = self.flat_grad_bf16().data_ptr() if self.mirror_bf16 and self.precision == "bf16" else None
        for k in range(S.BWD_GROUPS):
            ev = None if self.bwd_events is None else self.bwd_events[k]
            self._desc.bwd_events[k] = ev.value if hasattr(ev, "value") else ev
        self._grads_fresh = False
        with torch.cuda.device(x.device):
            check(lib.sea_temporal_backward(C.byref(self._desc), C.c_void_p(self._cache.data_ptr()),
                                            C.c_void_p(x.data_ptr()), C.c_void_p(ib.data_ptr()),
                                            C.c_void_p(dy.data_ptr()),
                                            None if dx is None else C.c_void_p(dx.data_ptr()), B, T,
                                            C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "temporal_backward")
        self.last_launches = lib.sea_last_launch_count()
        self.total_launches += self.last_launches
        return dx

    @torch.no_grad()
    def forward_nograd(self, x: torch.Tensor, ib: torch.Tensor, training: bool = False,
                       ws: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("sea_b200 has no CPU path: inputs must be CUDA tensors")
        if training and self.dropout > 0.0 and self.precision != "bf16":
            raise NotImplementedError("sea_b200: train-mode dropout runs in bf16 mode only")
        self._ensure(training)
        B, T, V, E = x.shape
        x = x.contiguous().float()
        auto_key = None
        inv = bool(self.ib_time_invariant) and not training
        if not inv and not training and self.auto_time_invariant and T > 1:
            inv, auto_key = self._auto_invariant(ib, B)
        ib = ib.contiguous().float()
        y = torch.empty_like(x)
        if ws is None:
            ws = self.workspace(B, T, training)
        self._desc.ib_time_invariant = int(inv)
        # train-mode nn.Dropout (attention probabilities, MLP and TIPI outputs): counter-based masks keyed by
        # a seed drawn from torch's CPU generator (torch.manual_seed makes runs repeatable); the backward of
        # this forward must see the same seed (autograd.py keeps it)
        self._desc.dropout_p = float(self.dropout) if training else 0.0
        if training and self.dropout > 0.0:
            self.last_dropout_seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        else:
            self.last_dropout_seed = 0
        self._desc.dropout_seed = self.last_dropout_seed
        self._desc.cond_cache, self._desc.cond_cache_bytes, self._desc.cond_cache_valid = None, 0, 0
        use_cc = inv and (self.cond_reuse or auto_key is not None) and T > 1
        if use_cc:
            key = (B, self._cache_key, auto_key)
            if self._cond_key != key or self._cond_buf is None:
                n = lib.sea_temporal_cond_cache_bytes(C.byref(self._desc), B)
                self._cond_buf = torch.empty(n, dtype=torch.uint8, device=self._dev)
                self._cond_key, self._cond_valid = key, False
            self._desc.cond_cache = self._cond_buf.data_ptr()
            self._desc.cond_cache_bytes = self._cond_buf.numel()
            self._desc.cond_cache_valid = int(self._cond_valid)
        with torch.cuda.device(x.device):
            check(lib.sea_temporal_forward(C.byref(self._desc), C.c_void_p(self._cache.data_ptr()),
                                           C.c_void_p(x.data_ptr()), C.c_void_p(ib.data_ptr()),
                                           C.c_void_p(y.data_ptr()), B, T, C.c_void_p(ws.data_ptr()),
                                           C.c_size_t(ws.numel()), int(training),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "temporal_forward")
        if use_cc:
            self._cond_valid = True
        self.last_launches = lib.sea_last_launch_count()
        self.total_launches += self.last_launches
        return y

    # The reference's rollout loops (utils/train_utils.py:171-175, 203-207) call model(seq, ib[:, :i+1]) with growing views of
    # ONE condition tensor.  The BASE tensor is tested once for time-invariance (one device sync per distinct tensor /
    # version); every later call on a view of it then takes the per-trajectory condition path and reuses the cached AdaLN /
    # TIPI condition rows, which depend on ib and the weights only — what rollout() is told explicitly.
    auto_time_invariant = os.environ.get("SEA_AUTO_INVARIANT", "1") != "0"

    def _auto_invariant(self, ib: torch.Tensor, B: int):
        base = ib._base if ib._base is not None else ib
        if base.dim() != 3 or not base.is_cuda or ib.dim() != 3:
            return False, None
        # identity of the base tensor OBJECT (weak reference) + its version counter — never its address: the caching
        # allocator hands the next batch's condition tensor the same address, and that one holds other values
        ref = getattr(self, "_ib_auto_ref", None)
        if ref is None or ref() is not base or self._ib_auto_version != base._version:
            self._ib_auto = bool((base == base[:, :1]).all().item())
            self._ib_auto_ref, self._ib_auto_version = weakref.ref(base), base._version
            self._ib_auto_gen = getattr(self, "_ib_auto_gen", 0) + 1
        if not self._ib_auto:
            return False, None
        # the cached condition rows belong to THESE trajectories of THIS tensor: generation, view origin, batch
        return True, (self._ib_auto_gen, ib.data_ptr() - base.data_ptr(), B)

    @torch.no_grad()
    def forward_into(self, x: torch.Tensor, ib: torch.Tensor, y: torch.Tensor, ws: torch.Tensor, *,
                     time_invariant: bool, cond_buf: Optional[torch.Tensor] = None,
                     cond_valid: bool = False) -> int:
        """Allocation-free inference forward (CUDA-graph capturable): fp32 x [B,T,V,E] — contiguous, or
        a prefix view ``seq[:, :T]`` of a longer [B,S,V,E] buffer (only the batch stride may differ) —
        ib [B,T,ib_num] -> y, activations in the caller's workspace ``ws``; the caller has already
        run ``_ensure(False)``.  Returns the number of kernels enqueued."""
        B, T, V, E = x.shape
        if x.dtype != torch.float32 or x.stride(3) != 1 or x.stride(2) != E or x.stride(1) != V * E:
            raise RuntimeError("sea_b200 forward_into: x must be fp32 [B,T,V,E] with contiguous trajectories")
        strided = x.stride(0) != T * V * E
        d = self._desc
        d.ib_time_invariant = int(time_invariant)
        if cond_buf is not None and time_invariant and T > 1:
            d.cond_cache, d.cond_cache_bytes, d.cond_cache_valid = cond_buf.data_ptr(), cond_buf.numel(), int(cond_valid)
        else:
            d.cond_cache, d.cond_cache_bytes, d.cond_cache_valid = None, 0, 0
        with torch.cuda.device(x.device):
            if strided:
                check(lib.sea_temporal_forward_strided(C.byref(d), C.c_void_p(self._cache.data_ptr()),
                                                       C.c_void_p(x.data_ptr()), C.c_int64(x.stride(0)),
                                                       C.c_void_p(ib.data_ptr()), C.c_void_p(y.data_ptr()), B, T,
                                                       C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "temporal_forward_strided")
            else:
                check(lib.sea_temporal_forward(C.byref(d), C.c_void_p(self._cache.data_ptr()),
                                               C.c_void_p(x.data_ptr()), C.c_void_p(ib.data_ptr()),
                                               C.c_void_p(y.data_ptr()), B, T, C.c_void_p(ws.data_ptr()),
                                               C.c_size_t(ws.numel()), 0,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "temporal_forward")
        return int(lib.sea_last_launch_count())

    @torch.no_grad()
    def step_into(self, x_t: torch.Tensor, ib_t: torch.Tensor, y_t: torch.Tensor, pos: int,
                  kv: torch.Tensor, max_len: int, ws: torch.Tensor, *, time_invariant: bool,
                  cond_buf: Optional[torch.Tensor] = None, cond_valid: bool = False) -> int:
        """KV-cached incremental step (sea_temporal_step): x_t [B,V,E] (new token at position ``pos``),
        ib_t [B,ib_num], y_t [B,V,E]; all three may be slices of [B,T,...] buffers (only the batch
        stride may be non-contiguous).  Allocation-free / graph-capturable; caller ran ``_ensure(False)``."""
        B = x_t.shape[0]
        for t_ in (x_t, y_t):
            if t_.dtype != torch.float32 or t_.stride(-1) != 1 or t_.stride(-2) != t_.shape[-1]:
                raise RuntimeError("sea_b200 step: x_t / y_t must be fp32 with contiguous [V,E] rows")
        d = self._desc
        d.ib_time_invariant = int(time_invariant)
        if cond_buf is not None and time_invariant:
            d.cond_cache, d.cond_cache_bytes, d.cond_cache_valid = cond_buf.data_ptr(), cond_buf.numel(), int(cond_valid)
        else:
            d.cond_cache, d.cond_cache_bytes, d.cond_cache_valid = None, 0, 0
        with torch.cuda.device(x_t.device):
            check(lib.sea_temporal_step(C.byref(d), C.c_void_p(self._cache.data_ptr()), C.c_void_p(kv.data_ptr()),
                                        C.c_size_t(kv.numel()), int(max_len),
                                        C.c_void_p(x_t.data_ptr()), C.c_int64(x_t.stride(0)),
                                        C.c_void_p(ib_t.data_ptr()), C.c_int64(ib_t.stride(0)),
                                        C.c_void_p(y_t.data_ptr()), C.c_int64(y_t.stride(0)), B, int(pos),
                                        C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "temporal_step")
        return int(lib.sea_last_launch_count())

    def __call__(self, x, ib):
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.module.parameters()))
        if needs_grad:
            from .autograd import temporal_apply
            return temporal_apply(self, x, ib)
        # nn.Dropout follows module.training, not the grad mode (models/base_blocks.py:47,194): a train-mode
        # module under torch.no_grad() still drops
        return self.forward_nograd(x, ib, training=bool(self.module.training) and self.dropout > 0.0)


def accelerate(model: nn.Module, precision: str = "bf16") -> nn.Module:
    """Rebind ``forward`` of an UNCHANGED reference ``TemporalModel`` instance (or the mirror) to
    the CUDA path.  Parameters are read in place; state_dict / optimizer objects keep working.

    The configuration both reference configs select (exchange_mode='sea', ib 'mlp' + 'add', add_info_after_cross) runs
    through the fused whole-model executor.  Every other block / ib variant the reference ships (pool / addition / simple
    exchange, fourier / linear ib layers, concat / attention / none addition: models/temporal.py:103-120, 197-312) keeps
    the reference's own block orchestration and gets the module-level CUDA path (sea_b200.modules: GEMMs, fused
    attention, norms and the MLP core as autograd Functions over the same kernels, bf16 mode)."""
    try:
        _check_modes(getattr(model, "exchange_mode", "sea"), getattr(model, "ib_scale_mode", "mlp"),
                     getattr(model, "ib_addition_mode", "add"),
                     getattr(model.blocks[0], "add_info_after_cross", True), getattr(model, "LN_type", "ln"))
        if getattr(model.blocks[0], "ib_mlp_layers", 1) not in (None, 1):
            raise NotImplementedError("ib_mlp_layers")
    except NotImplementedError:
        if precision != "bf16":
            raise NotImplementedError("non-default exchange / ib modes are accelerated module by module in bf16 mode only")
        from .modules import accelerate_modules
        accelerate_modules(model)
        return model
    drop = getattr(model.blocks[0], "dropout", 0.0)
    drop = float(getattr(drop, "p", drop))
    eng = TemporalEngine(model, precision=precision, dropout=drop)

    def forward(self, x, x_additional_info):
        assert x.shape[2] == self.num_variables, \
            f"Expected {self.num_variables} variables, but got {x.shape[2]}"
        eng.dropout = drop if self.training else 0.0
        return eng(x, x_additional_info)

    model.forward = types.MethodType(forward, model)
    model._sea_engine = eng
    return model

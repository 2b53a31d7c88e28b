"""Drop-in boundary checks that need no GPU: state_dict compatibility with the reference,
constructor/error behaviour, C-ABI symbol export."""
import ctypes
import json
import os
import re

import pytest
import torch

from tests.helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _manifest():
    with open(os.path.join(GOLDEN, "state_dict_manifest.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("tag,V,ln", [("small_adaln", 2, "adaln"), ("small_ln", 2, "ln"), ("small_v3", 3, "ln")])
def test_temporal_state_dict_matches_reference(tag, V, ln):
    from sea_b200.temporal import TemporalModel
    m = TemporalModel(1, 128, 2, 64, 2, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    ours = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    assert ours == _manifest()["temporal_" + tag]


def test_spatial_state_dict_matches_reference():
    from sea_b200.spatial import SpatialModel
    m = SpatialModel([[0, 1], [2]], 16, 48, 2, 8, 8, 2024, 0, 0.0, False)
    ours = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    assert ours == _manifest()["spatial_small"]
    with pytest.raises(NotImplementedError):
        SpatialModel([[0, 1], [2]], 16, 48, 2, 8, 8, 2024, 0, 0.0, True)


def test_temporal_init_distribution():
    from sea_b200.temporal import TemporalModel
    torch.manual_seed(0)
    m = TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln")
    sd = m.state_dict()
    w = sd["blocks.0.mlp.0.layers.0.weight"]
    assert abs(w.std().item() - 0.02) < 2e-3 and sd["blocks.0.mlp.0.layers.0.bias"].abs().max() == 0
    assert torch.all(sd["ln.0.weight"] == 1) and torch.all(sd["ln.0.bias"] == 0)
    assert abs(sd["ln.0.cond_mlp.2.weight"].std().item() - 0.02) < 2e-3
    assert torch.all(sd["blocks.0.mlp.0.layers.1.weight"] == 1)


def test_unsupported_modes_raise():
    from sea_b200.temporal import TemporalModel
    with pytest.raises(NotImplementedError):
        TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, "pool", "learnable", "mlp", "add", 1, 1, True, "ln")
    with pytest.raises(NotImplementedError):
        TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "fourier", "add", 1, 1, True, "ln")
    with pytest.raises(ValueError):
        TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "rms")


def test_no_cpu_fallback():
    from sea_b200.temporal import TemporalModel
    m = TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "ln")
    with pytest.raises(RuntimeError, match="no CPU path"):
        with torch.no_grad():
            m(torch.zeros(1, 4, 2, 128), torch.zeros(1, 4, 1))


def test_library_exports_every_declared_symbol():
    from sea_b200._lib import LIB_PATH
    hdr = open(os.path.join(ROOT, "include", "sea_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sea_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 10
    dll = ctypes.CDLL(LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(dll, n)]
    assert not missing, missing
    dll.sea_strerror.restype = ctypes.c_char_p
    assert b"unsupported" in dll.sea_strerror(-2).lower() or b"supported" in dll.sea_strerror(-2)
    assert dll.sea_version() >= 100


def test_ctypes_struct_sizes_match_header():
    """sizeof of every ctypes mirror == sizeof of the C struct (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from sea_b200 import _lib, _structs as S
    pairs = {"sea_gemm_epilogue": _lib.GemmEpilogue, "sea_gemm_problem": _lib.GemmProblem,
             "sea_norm_args": S.NormArgs, "sea_ln_gelu_args": S.LnGeluArgs, "sea_pack_args": S.PackArgs,
             "sea_attn_args": S.AttnArgs, "sea_param": S.Param, "sea_norm_params": S.NormParams,
             "sea_attn_params": S.AttnParams, "sea_stream_params": S.StreamParams,
             "sea_block_params": S.BlockParams, "sea_temporal_desc": S.TemporalDesc,
             "sea_norm_bwd_args": S.NormBwdArgs, "sea_ln_gelu_bwd_args": S.LnGeluBwdArgs,
             "sea_attn_bwd_args": S.AttnBwdArgs, "sea_tipi_bwd_args": S.TipiBwdArgs,
             "sea_spatial_layer": S.SpatialLayer, "sea_spatial_desc": S.SpatialDesc,
             "sea_adamw_hyper": __import__("sea_b200.optim", fromlist=["AdamWHyper"]).AdamWHyper,
             "sea_field_scaler": __import__("sea_b200.pipeline", fromlist=["FieldScaler"]).FieldScaler}
    src = '#include <stdio.h>\n#include "sea_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in pairs) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "s.c")
        open(cfile, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = dict(zip(out[::2], map(int, out[1::2])))
    for n, cls in pairs.items():
        assert ctypes.sizeof(cls) == sizes[n], (n, ctypes.sizeof(cls), sizes[n])
    from sea_b200.optim import _CHUNK_DT
    assert _CHUNK_DT.itemsize == 48  # sea_adamw_chunk: 5 pointers + 2 int32

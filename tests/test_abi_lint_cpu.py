"""Static checks of the drop-in boundary that need no GPU: every ctypes mirror lays its fields out exactly as the C struct
of include/sea_b200.h does (offsetof, not only sizeof), and every call site of a library entry point in the Python host
code passes the declared number of arguments with 64-bit scalars wrapped (a bare Python int is passed as a C int and
silently masked to 32 bits when no argtypes are set)."""
import ast
import ctypes
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sea_b200.h")


def _mirrors():
    from sea_b200 import _lib, _structs as S
    from sea_b200.optim import AdamWHyper
    from sea_b200.pipeline import FieldScaler
    from sea_b200.rollout import ProfileSummary
    return {"sea_gemm_epilogue": _lib.GemmEpilogue, "sea_gemm_problem": _lib.GemmProblem,
            "sea_norm_args": S.NormArgs, "sea_ln_gelu_args": S.LnGeluArgs, "sea_pack_args": S.PackArgs,
            "sea_attn_args": S.AttnArgs, "sea_param": S.Param, "sea_norm_params": S.NormParams,
            "sea_attn_params": S.AttnParams, "sea_stream_params": S.StreamParams,
            "sea_block_params": S.BlockParams, "sea_temporal_desc": S.TemporalDesc,
            "sea_norm_bwd_args": S.NormBwdArgs, "sea_ln_gelu_bwd_args": S.LnGeluBwdArgs,
            "sea_attn_bwd_args": S.AttnBwdArgs, "sea_tipi_bwd_args": S.TipiBwdArgs,
            "sea_spatial_layer": S.SpatialLayer, "sea_spatial_desc": S.SpatialDesc,
            "sea_adamw_hyper": AdamWHyper, "sea_field_scaler": FieldScaler, "sea_profile_summary": ProfileSummary}


def test_ctypes_field_offsets_match_header():
    pairs = _mirrors()
    lines = [f'printf("{n}.{f[0]} %zu\\n", offsetof({n}, {f[0]}));' for n, cls in pairs.items() for f in cls._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "sea_b200.h"\nint main(){' + "\n".join(lines) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "o.c"), os.path.join(td, "o")
        with open(cfile, "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])   # same field NAMES, too
        out = subprocess.check_output([exe]).decode().split()
    offs = dict(zip(out[::2], map(int, out[1::2])))
    assert len(offs) > 250
    wrong = [(n, f[0], getattr(cls, f[0]).offset, offs[f"{n}.{f[0]}"]) for n, cls in pairs.items() for f in cls._fields_
             if getattr(cls, f[0]).offset != offs[f"{n}.{f[0]}"]]
    assert not wrong, wrong


def _declarations():
    hdr = open(HEADER).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    decl = {}
    for m in re.finditer(r"\b(?:int|size_t|void|const char\*)\s+(sea_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return decl


def _python_sources():
    for top in ("sea_b200", "tests", "scripts"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    yield os.path.join(dp, f)
    yield os.path.join(ROOT, "bench.py")
    yield os.path.join(ROOT, "__graft_entry__.py")


def test_every_ffi_call_site_matches_its_declaration():
    decl = _declarations()
    assert len(decl) >= 70
    sites, bad = 0, []
    wrap64 = re.compile(r"^(C|ctypes)\.c_(int64|uint64|size_t|longlong|ulonglong)\(")
    for path in _python_sources():
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in decl):
                continue
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            params = decl[node.func.attr]
            sites += 1
            where = f"{os.path.relpath(path, ROOT)}:{node.lineno} {node.func.attr}"
            if len(node.args) != len(params) or node.keywords:
                bad.append(f"{where}: {len(node.args)} arguments, declared {len(params)}")
                continue
            for a, prm in zip(node.args, params):
                src = ast.unparse(a)
                is_ptr = "*" in prm or "sea_stream_t" in prm
                if not is_ptr and re.search(r"\b(int64_t|uint64_t|size_t)\b", prm) and not wrap64.match(src):
                    bad.append(f"{where}: `{src}` for `{prm}` is not wrapped in a 64-bit ctypes scalar")
                if is_ptr and re.fullmatch(r".*\.data_ptr\(\)|-?[1-9]\d*", src):
                    bad.append(f"{where}: `{src}` for `{prm}` is a bare Python int (would be passed as a 32-bit C int)")
    assert sites >= 100, sites
    assert not bad, "\n".join(bad)


def test_size_returning_entry_points_have_a_size_t_restype():
    """ctypes assumes `int` results: every size_t-returning entry point must be given its restype in _lib.py."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    size_fns = set(re.findall(r"\bsize_t\s+(sea_[a-z0-9_]+)\s*\(", hdr))
    assert size_fns
    from sea_b200._lib import lib
    missing = [n for n in sorted(size_fns) if getattr(lib, n).restype is not ctypes.c_size_t]
    assert not missing, missing

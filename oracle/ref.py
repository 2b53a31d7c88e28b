"""ORACLE SUPPORT (test infrastructure, never imported by sea_b200/) — the UNMODIFIED reference.

The reference is pure Python, so "building" it is a file copy: ``stage()`` (called by
``__graft_entry__.build()`` in the build container, where /root/reference exists) copies the reference's
own packages byte-for-byte into ``oracle/_ref/`` — git-ignored, so no reference source enters the history,
but NOT gpurun-ignored, so the copy travels to the GPU box like a built ``.so`` — and writes a sha256
manifest beside it.  ``load()`` puts that directory on ``sys.path`` (falling back to /root/reference when
the staged copy is absent), stubs the plotting / logging packages the reference imports at module level but
that are not installed (matplotlib, h5py, wandb: utils/modular_testing.py:4-5, train/train_encoder.py:4),
and returns the reference's modules.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
may call this.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("SEA_REFERENCE", "/root/reference")
PACKAGES = ("models", "utils", "train", "configs")
FILES = ("main.py", "__init__.py", "LICENSE")
STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.tri", "h5py", "wandb")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(force: bool = False) -> str | None:
    """Copy the reference's python packages into oracle/_ref (no edits).  Returns the staged path, or None
    when there is no reference tree to copy from (the GPU box: it received the staged copy already)."""
    if not os.path.isdir(SOURCE):
        return STAGED if os.path.isdir(STAGED) else None
    manifest_path = os.path.join(STAGED, "MANIFEST.json")
    files = {}
    for pkg in PACKAGES:
        for dirpath, _, names in os.walk(os.path.join(SOURCE, pkg)):
            for n in names:
                if n.endswith(".py"):
                    full = os.path.join(dirpath, n)
                    files[os.path.relpath(full, SOURCE)] = _sha(full)
    for n in FILES:
        if os.path.exists(os.path.join(SOURCE, n)):
            files[n] = _sha(os.path.join(SOURCE, n))
    if not force and os.path.exists(manifest_path):
        try:
            with open(manifest_path) as f:
                if json.load(f).get("files") == files and all(
                        os.path.exists(os.path.join(STAGED, r)) for r in files):
                    return STAGED
        except (OSError, ValueError):
            pass
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    for rel in files:
        dst = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, rel), dst)
    with open(manifest_path, "w") as f:
        json.dump({"source": SOURCE, "files": files}, f, indent=1, sort_keys=True)
    return STAGED


def verify() -> bool:
    """True when every staged file still has the sha256 recorded at staging time (i.e. is unmodified)."""
    try:
        with open(os.path.join(STAGED, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
    except (OSError, ValueError, KeyError):
        return False
    return all(os.path.exists(os.path.join(STAGED, r)) and _sha(os.path.join(STAGED, r)) == h
               for r, h in files.items())


def root() -> str | None:
    if os.path.isdir(os.path.join(STAGED, "models")):
        return STAGED
    if os.path.isdir(os.path.join(SOURCE, "models")):
        return SOURCE
    return None


def available() -> bool:
    return root() is not None


class _Anything:
    """Attribute sink for the plotting stubs: any name resolves, calling it is an error we want to see."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything(f"{self._name}.{k}")

    def __call__(self, *a, **kw):
        raise RuntimeError(f"{self._name} is a stub (plotting / logging is outside the hot path and not installed)")


def _install_stubs():
    for name in STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)

        def _attr(k, _n=name):
            if k.startswith("__"):       # inspect / importlib probe dunders (__file__, __path__ ...): absent
                raise AttributeError(k)
            return _Anything(f"{_n}.{k}")
        mod.__getattr__ = _attr  # type: ignore[attr-defined]
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


def load() -> types.SimpleNamespace:
    """Import the reference (unmodified) and return its modules:
    .temporal .base_blocks .encoder_decoder .train_temporal .train_utils .data_processors .cfg_cylinder
    .cfg_multiphase, plus .root (where it was imported from)."""
    r = root()
    if r is None:
        raise RuntimeError("reference not available: neither oracle/_ref (run __graft_entry__.build() in the "
                           "build container) nor /root/reference exists")
    import torch  # noqa: F401  (before the stubs: torch's own import inspects sys.modules)
    if r not in sys.path:
        sys.path.insert(0, r)
    _install_stubs()
    ns = types.SimpleNamespace(root=r)
    ns.base_blocks = importlib.import_module("models.base_blocks")
    ns.temporal = importlib.import_module("models.temporal")
    ns.encoder_decoder = importlib.import_module("models.encoder_decoder")
    ns.cfg_cylinder = importlib.import_module("configs.cylinder_flow")
    ns.cfg_multiphase = importlib.import_module("configs.multiphase_flow")
    ns.train_utils = importlib.import_module("utils.train_utils")
    ns.data_processors = importlib.import_module("utils.data_processors")
    ns.train_temporal = importlib.import_module("train.train_temporal")
    for m in (ns.base_blocks, ns.temporal, ns.encoder_decoder, ns.train_utils, ns.train_temporal):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(r)), (m.__file__, r)
    return ns


def temporal_config(name: str) -> dict:
    """The reference's own temporal config dict (configs/<name>.py:get_config_temporal), device left to the caller."""
    ns = load()
    mod = ns.cfg_cylinder if name == "cylinder_flow" else ns.cfg_multiphase
    cfg = dict(mod.get_config_temporal())
    cfg["use_wandb"] = False
    return cfg


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv), "verified" if verify() else "NOT verified")

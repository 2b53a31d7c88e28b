"""The one-line hook that switches the reference's UNCHANGED entry points to the CUDA path.

    import sea_b200; sea_b200.install()        # in a launcher / sitecustomize, before main.py runs

``main.py``, ``train/train_temporal.py``, ``utils/*.py`` and ``configs/*.py`` stay byte-identical: the hook
wraps two factory functions of the reference so that the modules they return run through libsea_b200.so —

* ``train.train_temporal.get_model`` (train/train_temporal.py:190-223): the returned ``TemporalModel`` is
  ``accelerate()``d (its ``forward`` is rebound; parameters, ``state_dict`` and the ``torch.optim.AdamW`` built
  by ``initialize_optimizer`` keep working on the same ``nn.Parameter`` objects);
* ``utils.data_processors.ProcessData.initialize_spatial_model`` (utils/data_processors.py:305-317): the
  returned ``SpatialModel`` is ``accelerate_spatial()``d, so ``process_data`` / ``decode_data`` encode and
  decode through the fused codec kernels.

Models on a CPU device are returned untouched (sea_b200 has no CPU path and does not pretend to).
``uninstall()`` restores the originals.
"""
from __future__ import annotations

import importlib

_saved = {}


def _is_cuda(device) -> bool:
    import torch
    return torch.device(device).type == "cuda"


def install(precision: str = "bf16", fused_optimizer: bool = False, spatial: bool = True,
            spatial_precision: str = "fp32") -> None:
    """Wrap the reference factories (idempotent).  ``fused_optimizer=True`` additionally replaces the
    ``torch.optim.AdamW`` the reference builds (utils/train_utils.py:33-39) by ``sea_b200.optim.AdamW`` with the
    same hyper-parameters (same arithmetic, one launch per step)."""
    from .spatial import accelerate_spatial
    from .temporal import accelerate

    tt = importlib.import_module("train.train_temporal")
    if "get_model" not in _saved:
        ref_get_model = tt.get_model
        _saved["get_model"] = (tt, ref_get_model)

        def get_model(config, device):
            model, loss_fn, optimizer = ref_get_model(config, device)
            if not _is_cuda(device):
                return model, loss_fn, optimizer
            accelerate(model, precision=_saved["precision"])
            if _saved["fused_optimizer"] and not isinstance(optimizer, tuple):
                from .optim import AdamW
                g = optimizer.param_groups[0]
                optimizer = AdamW(model.parameters(), lr=g["lr"], betas=g["betas"], eps=g["eps"],
                                  weight_decay=g["weight_decay"], engine=model._sea_engine)
            return model, loss_fn, optimizer

        get_model.__wrapped__ = ref_get_model
        tt.get_model = get_model
    _saved["precision"], _saved["fused_optimizer"] = precision, bool(fused_optimizer)
    _saved["spatial_precision"] = spatial_precision

    if spatial and "init_spatial" not in _saved:
        dp = importlib.import_module("utils.data_processors")
        ref_init = dp.ProcessData.initialize_spatial_model
        _saved["init_spatial"] = (dp.ProcessData, ref_init)

        def initialize_spatial_model(self):
            model = ref_init(self)
            if _is_cuda(self.device) and not getattr(model, "variational", False):
                accelerate_spatial(model, precision=_saved.get("spatial_precision", "fp32"))
            return model

        initialize_spatial_model.__wrapped__ = ref_init
        dp.ProcessData.initialize_spatial_model = initialize_spatial_model


def uninstall() -> None:
    if "get_model" in _saved:
        mod, fn = _saved.pop("get_model")
        mod.get_model = fn
    if "init_spatial" in _saved:
        cls, fn = _saved.pop("init_spatial")
        cls.initialize_spatial_model = fn
    _saved.pop("precision", None)
    _saved.pop("fused_optimizer", None)
    _saved.pop("spatial_precision", None)

"""Resident fields -> latents -> rollout -> fields pipeline.

What the reference does around the temporal model for an evaluation rollout
(``utils/train_utils.py:199-230``, ``train/train_temporal.py:137-165``, ``utils/data_processors.py:330-363,
484-573``):

    scale fields per group (MinMaxScaler, CPU) -> DataPartitioner2D.create_partitions (CPU, 64 boolean masks)
    -> permute [S,P,C,F] -> [S,P,F,C] -> SpatialModel.encode in chunks, every chunk ``.cpu()`` -> torch.cat
    -> transform_processed_data (reshape / permute / reshape) -> ... rollout loop ...
    -> inverse_transform_processed_data -> ProcessData.decode_data (re-creates the SpatialModel, re-loads its
       checkpoint from disk, ``.cpu()``) -> permute -> inverse_partition + inverse scaling on the CPU -> ``.to(device)``

Here every stage stays in HBM and the glue is folded into the kernels on either side of it:

    sea_patch_gather_scaled   scaler.transform + patchify + the [S,P,C,F]->[S,P,F,C] permute, one pass
    sea_spatial_encode        latent_layout = 1: writes [tr*T, G, P*D] = transform_processed_data's result directly
    rollout()                 graphed prefix-recompute loop (or the KV-cached engine)
    sea_spatial_decode        latent_layout = 1: reads the temporal layout directly (inverse_transform folded)
    sea_patch_scatter_scaled  the permute back + inverse_partition + scaler.inverse_transform, one pass

The spatial codec stays resident (no per-call re-creation / checkpoint re-load).  Results are identical to the
reference chain: index work and the scaler arithmetic bit-exact, codec / temporal model to their parity bars
(tests/test_pipeline_gpu.py runs the reference's own functions and classes beside it on the same GPU).
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from ._lib import check, lib
from .patchify import DataPartitioner2D
from .rollout import rollout


class FieldScaler(C.Structure):
    """sea_field_scaler (include/sea_b200.h)."""
    _fields_ = [("min_val", C.c_float), ("max_val", C.c_float), ("lo", C.c_float), ("range", C.c_float),
                ("enabled", C.c_int32), ("reserved", C.c_int32)]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _codec_of(spatial_model):
    codec = getattr(spatial_model, "_sea_codec", None)          # accelerate_spatial() on a reference instance
    if codec is None and hasattr(spatial_model, "_codec"):
        codec = spatial_model._codec()                           # the mirror class
    if codec is None:
        raise RuntimeError("spatial model is not on the sea_b200 path: construct sea_b200.spatial.SpatialModel or "
                           "call accelerate_spatial() on the reference instance")
    return codec


class ResidentPipeline:
    """``fields[tr, T, n_cells, F]`` in, ``fields`` out, nothing leaves the device in between.

    temporal_model / spatial_model: sea_b200 mirrors or accelerate()d reference instances (frozen, eval).
    x_coords, y_coords: the mesh (``MeshProcessor.coordinates``).  field_groups: ``config['field_groups']``.
    feature_range: ``config['scale_feature_range']`` (None = no scaling).  m, n: patch grid (``config['m'], ['n']``)."""

    def __init__(self, temporal_model, spatial_model, x_coords, y_coords, field_groups: Sequence[Sequence[int]],
                 feature_range: Optional[Tuple[float, float]] = None, m: int = 9, n: int = 9, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sea_b200.pipeline has no CPU path")
        self.temporal, self.spatial = temporal_model, spatial_model
        self.codec = _codec_of(spatial_model)
        self.partitioner = DataPartitioner2D(x_coords, y_coords, m=m, n=n, pad_id=-1, pad_field_value=0, device=self.device)
        self.partitioner._build_index()
        self.field_groups = [list(g) for g in field_groups]
        self.n_fields = max(max(g) for g in self.field_groups) + 1
        self.feature_range = feature_range
        self.minmax = None                                         # per group (min, max) once fitted / loaded
        self._scalers = (FieldScaler * self.n_fields)()            # identity until fitted
        self.n_patches = (m - 1) * (n - 1)

    # ------------------------------------------------------------------ MinMaxScaler per field group
    def fit_scalers(self, fields: torch.Tensor) -> None:
        """MinMaxScaler.fit on ``fields[..., group]`` of the training snapshots (utils/data_processors.py:233-243,
        :492-494): global min / max per field group."""
        self.set_scalers([(torch.min(fields[..., g]).item(), torch.max(fields[..., g]).item()) for g in self.field_groups])

    def set_scalers(self, minmax: Sequence[Tuple[float, float]]) -> None:
        """(min_val, max_val) per field group, e.g. loaded from the reference's ``*_min_max_values.pt`` files."""
        if len(minmax) != len(self.field_groups):
            raise ValueError("one (min, max) pair per field group")
        self.minmax = [(float(a), float(b)) for a, b in minmax]
        if self.feature_range is None:
            return
        lo, hi = self.feature_range
        for (mn, mx), group in zip(self.minmax, self.field_groups):
            if mn == mx:
                raise ValueError("Data has zero variance")
            for f in group:
                self._scalers[f] = FieldScaler(mn, mx, float(lo), float(hi - lo), 1, 0)

    # ------------------------------------------------------------------ stages
    def patchify(self, fields: torch.Tensor) -> torch.Tensor:
        """[S, n_cells, F] raw fields -> scaled padded patches [S, P, F, C] (the SpatialModel input, SEA_isolate)."""
        fields = fields.to(self.device, torch.float32).contiguous()
        S_, N, F_ = fields.shape
        part = self.partitioner
        if N != part.x_coords.numel() or F_ != self.n_fields:
            raise ValueError(f"fields must be [S, {part.x_coords.numel()}, {self.n_fields}], got {tuple(fields.shape)}")
        P, Cc = part.index_map_tensor.shape
        out = torch.empty(S_, P, F_, Cc, dtype=torch.float32, device=self.device)
        for s0 in range(0, S_, 32768):                             # grid.z limit of the gather
            s1 = min(S_, s0 + 32768)
            with torch.cuda.device(self.device):
                check(lib.sea_patch_gather_scaled(C.c_void_p(fields[s0:s1].data_ptr()), N, F_, self._scalers,
                                                  C.c_void_p(part.index_map_tensor.data_ptr()), s1 - s0, P, Cc,
                                                  C.c_float(part.pad_field_value), 1, C.c_void_p(out[s0:s1].data_ptr()),
                                                  _stream()), "patch_gather_scaled")
        return out

    def unpatchify(self, patches: torch.Tensor) -> torch.Tensor:
        """[S, P, F, C] decoder output -> unscaled fields [S, n_cells, F] (inverse_scale_and_unpatch)."""
        patches = patches.to(self.device, torch.float32).contiguous()
        S_, P, F_, Cc = patches.shape
        part = self.partitioner
        N = part.x_coords.numel()
        out = torch.empty(S_, N, F_, dtype=torch.float32, device=self.device)
        for s0 in range(0, S_, 32768):
            s1 = min(S_, s0 + 32768)
            with torch.cuda.device(self.device):
                check(lib.sea_patch_scatter_scaled(C.c_void_p(patches[s0:s1].data_ptr()),
                                                   C.c_void_p(part.index_map_tensor.data_ptr()), s1 - s0, P, Cc, F_, N, 1,
                                                   self._scalers, C.c_void_p(out[s0:s1].data_ptr()), _stream()),
                      "patch_scatter_scaled")
        return out

    @torch.no_grad()
    def encode_fields(self, fields: torch.Tensor) -> torch.Tensor:
        """[tr, T, n_cells, F] -> temporal-model latents [tr, T, G, P*D] (= transform_processed_data(encode(...)))."""
        tr, T = fields.shape[:2]
        x = self.patchify(fields.reshape(tr * T, *fields.shape[2:]))
        if x.shape[-1] != self.codec_n_inp():
            raise ValueError(f"mesh has {x.shape[-1]} cells in its fullest patch but the encoder was built for "
                             f"n_inp = {self.codec_n_inp()}")
        z = self.codec.encode(x, fix_pad=False, latent_layout=1)   # [S, G, P*D]
        return z.view(tr, T, z.shape[1], z.shape[2])

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """[tr, T, G, P*D] -> fields [tr, T, n_cells, F]."""
        tr, T, G, PD = latents.shape
        out = self.codec.decode(latents.reshape(tr * T, G, PD), latent_layout=1)   # [S, P, F, C]
        rec = self.unpatchify(out)
        return rec.view(tr, T, rec.shape[1], rec.shape[2])

    def codec_n_inp(self) -> int:
        self.codec._ensure()
        return self.codec._dims["C"]

    @torch.no_grad()
    def rollout_fields(self, fields0: torch.Tensor, ib: torch.Tensor, steps: int, cached: bool = False,
                       return_latents: bool = False):
        """fields0 [tr, 1, n_cells, F] (the initial snapshot), ib [tr, >=steps, ib_num] -> predicted fields
        [tr, steps, n_cells, F]: encode, `steps` autoregressive steps of the temporal model (the loop of
        utils/train_utils.py:202-209), decode, unpatch, unscale — all on the device."""
        if fields0.dim() == 3:
            fields0 = fields0[:, None]
        z0 = self.encode_fields(fields0)
        lat = rollout(self.temporal, z0, ib.to(self.device, torch.float32), steps, cached=cached)
        rec = self.decode_latents(lat)
        return (rec, lat) if return_latents else rec

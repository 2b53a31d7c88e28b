"""Host-side mirrors of the reference's ``DataPartitioner2D`` / ``DataPartitioner3D`` (utils/data_processors.py:9-223) backed
by the CUDA kernels of ``csrc/patchify.cu``.

Same constructor arguments, ``create_partitions(vars)`` / ``inverse_partition(external_partitions,
time_dim)`` with the reference's return structure (a list of ``(coords [C,2], fields [S,C,F])`` per
patch and the padded index map) — the list entries are views into one stacked ``[S, P, C, F]``
tensor (``stacked_fields``), which is what ``patchify_and_scale`` builds from them (:529), so no
per-patch allocations happen.  ``stacked_fields_pfc()`` returns the ``[S, P, F, C]`` layout the
SpatialModel consumes.  torch supplies memory and the reference's own boundary recipe
(``torch.min/max/linspace`` on the coordinates, :27-31); bucketize, compaction, gather and scatter are
the library's kernels.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DataPartitioner2D:
    def __init__(self, x_coords, y_coords, m=9, n=9, pad_id=-1, pad_field_value=0, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sea_b200.patchify has no CPU path: pass a CUDA device")
        if pad_id >= 0:
            raise ValueError("pad_id must be negative")
        self.x_coords = x_coords.to(self.device).float().contiguous()
        self.y_coords = y_coords.to(self.device).float().contiguous()
        self.full_coords = torch.stack((self.x_coords, self.y_coords), dim=1)
        self.m, self.n = int(m), int(n)
        self.pad_id, self.pad_field_value = int(pad_id), float(pad_field_value)
        self.index_map_tensor: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ index map (mesh is static)
    def _build_index(self):
        N, P = self.x_coords.numel(), (self.m - 1) * (self.n - 1)
        # utils/data_processors.py:27-31 verbatim (host plumbing: 2 x m numbers)
        x_min, x_max = torch.min(self.x_coords), torch.max(self.x_coords)
        y_min, y_max = torch.min(self.y_coords), torch.max(self.y_coords)
        xb = torch.linspace(x_min, x_max, self.m, device=self.device).float().contiguous()
        yb = torch.linspace(y_min, y_max, self.n, device=self.device).float().contiguous()
        self.patch_id = torch.empty(N, dtype=torch.int32, device=self.device)
        self.counts = torch.empty(P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.sea_patch_bucketize(C.c_void_p(self.x_coords.data_ptr()), C.c_void_p(self.y_coords.data_ptr()),
                                          N, C.c_void_p(xb.data_ptr()), self.m, C.c_void_p(yb.data_ptr()), self.n,
                                          C.c_void_p(self.patch_id.data_ptr()), C.c_void_p(self.counts.data_ptr()),
                                          _stream()), "patch_bucketize")
            self.capacity = int(self.counts.max().item())     # the one data-dependent size (max cells per patch)
            self.index_map_tensor = torch.empty(P, self.capacity, dtype=torch.int64, device=self.device)
            check(lib.sea_patch_index_map(C.c_void_p(self.patch_id.data_ptr()), N, P, self.capacity,
                                          C.c_int64(self.pad_id), C.c_void_p(self.index_map_tensor.data_ptr()),
                                          _stream()), "patch_index_map")
        # padded coordinates per patch (pad_field_value where padded), :69-71
        valid = self.index_map_tensor >= 0
        safe = self.index_map_tensor.clamp_min(0)
        self.stacked_coords = torch.where(valid[..., None], self.full_coords[safe],
                                          torch.full((), self.pad_field_value, device=self.device))

    # ------------------------------------------------------------------ reference API
    def gather(self, vars: Sequence[torch.Tensor], layout_pfc: bool = False) -> torch.Tensor:
        """vars: F tensors [S, n_cells] -> [S, P, C, F] (or [S, P, F, C])."""
        if self.index_map_tensor is None:
            self._build_index()
        var_list = [v.to(self.device).float() for v in vars if v is not None]
        if len(var_list) == 0:
            raise ValueError("At least one variable must be provided")
        stacked = torch.stack(var_list, dim=0).contiguous()          # [F, S, N]: one pitch for all fields
        F_, S, N = stacked.shape
        P, Cc = self.index_map_tensor.shape
        out = torch.empty((S, P, F_, Cc) if layout_pfc else (S, P, Cc, F_), dtype=torch.float32, device=self.device)
        ptrs = (C.c_void_p * F_)(*[C.c_void_p(stacked[f].data_ptr()) for f in range(F_)])
        with torch.cuda.device(self.device):
            check(lib.sea_patch_gather(ptrs, F_, C.c_int64(N), C.c_void_p(self.index_map_tensor.data_ptr()), S, P, Cc,
                                       C.c_float(self.pad_field_value), int(layout_pfc), C.c_void_p(out.data_ptr()),
                                       _stream()), "patch_gather")
        self._n_fields = F_
        return out

    def create_partitions(self, vars) -> Tuple[List[Tuple[torch.Tensor, torch.Tensor]], List[torch.Tensor]]:
        self.stacked_fields = self.gather(vars)                       # [S, P, C, F]
        P = self.index_map_tensor.shape[0]
        self.padded_partitions = [(self.stacked_coords[p], self.stacked_fields[:, p]) for p in range(P)]
        self.padded_index_map = [self.index_map_tensor[p] for p in range(P)]
        return self.padded_partitions, self.padded_index_map

    def scatter(self, part: torch.Tensor, layout_pfc: bool = False) -> torch.Tensor:
        """[S, P, C, F] (or [S, P, F, C]) -> [S, n_cells, F]; every cell belongs to exactly one patch."""
        part = part.to(self.device).float().contiguous()
        S, P = part.shape[0], part.shape[1]
        Cc, F_ = (part.shape[3], part.shape[2]) if layout_pfc else (part.shape[2], part.shape[3])
        N = self.x_coords.numel()
        out = torch.empty(S, N, F_, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.sea_patch_scatter(C.c_void_p(part.data_ptr()), C.c_void_p(self.index_map_tensor.data_ptr()), S, P,
                                        Cc, F_, N, int(layout_pfc), C.c_void_p(out.data_ptr()), _stream()),
                  "patch_scatter")
        return out

    def inverse_partition(self, external_partitions=None, time_dim=None):
        partitions = external_partitions if external_partitions is not None else self.padded_partitions
        fields = torch.stack([f for _, f in partitions], dim=1)       # [S, P, C, F]
        rec = self.scatter(fields)
        if time_dim is not None:
            rec = rec[:time_dim]
        # coordinates: scatter of the padded coordinates = the mesh itself (:106-107)
        return self.full_coords.clone(), rec


class DataPartitioner3D(DataPartitioner2D):
    """Mirror of the reference's ``DataPartitioner3D`` (utils/data_processors.py:114-223): same constructor (the variables
    are given to the constructor, ``create_partitions()`` takes none), (m-1)(n-1)(k-1) patches ordered x-major, then y,
    then z.  Shares the compaction / gather / scatter kernels with the 2-D partitioner; only the bucketize differs."""

    def __init__(self, x_coords, y_coords, z_coords, vars, m=9, n=9, k=9, pad_id=-1, pad_field_value=0, device="cuda"):
        super().__init__(x_coords, y_coords, m=m, n=n, pad_id=pad_id, pad_field_value=pad_field_value, device=device)
        self.z_coords = z_coords.to(self.device).float().contiguous()
        self.full_coords = torch.stack((self.x_coords, self.y_coords, self.z_coords), dim=1)
        self.k = int(k)
        self.var_list = [v.to(self.device).float() for v in vars if v is not None]
        if len(self.var_list) == 0:
            raise ValueError("At least one variable must be provided")

    def _build_index(self):
        N, P = self.x_coords.numel(), (self.m - 1) * (self.n - 1) * (self.k - 1)
        bounds = []
        for coords, steps in ((self.x_coords, self.m), (self.y_coords, self.n), (self.z_coords, self.k)):   # :133-139
            bounds.append(torch.linspace(torch.min(coords), torch.max(coords), steps, device=self.device).float().contiguous())
        xb, yb, zb = bounds
        self.patch_id = torch.empty(N, dtype=torch.int32, device=self.device)
        self.counts = torch.empty(P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.sea_patch_bucketize3d(C.c_void_p(self.x_coords.data_ptr()), C.c_void_p(self.y_coords.data_ptr()),
                                            C.c_void_p(self.z_coords.data_ptr()), N, C.c_void_p(xb.data_ptr()), self.m,
                                            C.c_void_p(yb.data_ptr()), self.n, C.c_void_p(zb.data_ptr()), self.k,
                                            C.c_void_p(self.patch_id.data_ptr()), C.c_void_p(self.counts.data_ptr()),
                                            _stream()), "patch_bucketize3d")
            self.capacity = int(self.counts.max().item())
            self.index_map_tensor = torch.empty(P, self.capacity, dtype=torch.int64, device=self.device)
            check(lib.sea_patch_index_map(C.c_void_p(self.patch_id.data_ptr()), N, P, self.capacity,
                                          C.c_int64(self.pad_id), C.c_void_p(self.index_map_tensor.data_ptr()),
                                          _stream()), "patch_index_map")
        valid = self.index_map_tensor >= 0
        safe = self.index_map_tensor.clamp_min(0)
        self.stacked_coords = torch.where(valid[..., None], self.full_coords[safe],
                                          torch.full((), self.pad_field_value, device=self.device))

    def create_partitions(self):
        return super().create_partitions(self.var_list)

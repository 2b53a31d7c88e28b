"""Build-container only: golden vectors of the UNMODIFIED reference DataPartitioner2D
(/root/reference/utils/data_processors.py) -> tests/golden/patchify_small.npz.
matplotlib / h5py are not installed; utils.modular_testing only needs them for plotting, so empty
stub modules are placed in sys.modules before the import."""
import os
import sys
import types

import numpy as np
import torch

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "h5py"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference")
from utils.data_processors import DataPartitioner2D, DataPartitioner3D  # noqa: E402

rng = np.random.RandomState(20241018)
cases = {}
for tag, N, S, F, cluster in (("uniform", 700, 3, 3, False), ("clustered", 300, 2, 2, True)):
    x = rng.rand(N).astype(np.float32) * 2.2 - 0.2
    y = rng.rand(N).astype(np.float32)
    if cluster:          # leave many patches empty, put cells exactly on boundaries
        x = (np.round(x * 3) / 3).astype(np.float32)
        y = np.where(y > 0.5, y, 0.25 * y).astype(np.float32)
    vars_ = [rng.randn(S, N).astype(np.float32) for _ in range(F)]
    part = DataPartitioner2D(torch.from_numpy(x), torch.from_numpy(y), m=9, n=9, pad_id=-1, pad_field_value=0)
    padded, imap = part.create_partitions([torch.from_numpy(v) for v in vars_])
    fields = torch.stack([p[1] for p in padded], dim=1).numpy()          # [S, P, C, F]
    coords = torch.stack([p[0] for p in padded], dim=0).numpy()          # [P, C, 2]
    rc, rf = part.inverse_partition(padded)
    cases[tag] = dict(x=x, y=y, vars=np.stack(vars_, 0), index_map=torch.stack(imap, 0).numpy(), fields=fields,
                      coords=coords, recon=rf.numpy())
    assert np.array_equal(rf.numpy(), np.stack(vars_, 2))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "patchify_small.npz")
np.savez_compressed(out, **{f"{t}_{k}": v for t, c in cases.items() for k, v in c.items()})
print("wrote", out, {t: c["index_map"].shape for t, c in cases.items()})


# ---- DataPartitioner3D (utils/data_processors.py:114-223) -> tests/golden/patchify3d_small.npz
cases3 = {}
for tag, N, S, F, (m, n, k), cluster in (("uniform", 900, 2, 3, (5, 4, 6), False), ("clustered", 400, 3, 2, (9, 9, 9), True)):
    x = rng.rand(N).astype(np.float32) * 2.2 - 0.2
    y = rng.rand(N).astype(np.float32)
    z = (rng.rand(N).astype(np.float32) - 0.5) * 0.7
    if cluster:
        x = (np.round(x * 3) / 3).astype(np.float32)
        z = np.where(z > 0, z, 0.25 * z).astype(np.float32)
    vars_ = [rng.randn(S, N).astype(np.float32) for _ in range(F)]
    part = DataPartitioner3D(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(z), [torch.from_numpy(v) for v in vars_],
                             m=m, n=n, k=k, pad_id=-1, pad_field_value=0)
    padded, imap = part.create_partitions()
    fields = torch.stack([p[1] for p in padded], dim=1).numpy()
    coords = torch.stack([p[0] for p in padded], dim=0).numpy()
    rc, rf = part.inverse_partition(padded)
    cases3[tag] = dict(x=x, y=y, z=z, vars=np.stack(vars_, 0), mnk=np.array([m, n, k]), index_map=torch.stack(imap, 0).numpy(),
                       fields=fields, coords=coords, recon=rf.numpy())
    assert np.array_equal(rf.numpy(), np.stack(vars_, 2))
out3 = os.path.join(os.path.dirname(out), "patchify3d_small.npz")
np.savez_compressed(out3, **{f"{t}_{k}": v for t, c in cases3.items() for k, v in c.items()})
print("wrote", out3, {t: c["index_map"].shape for t, c in cases3.items()})

"""GPU patchify kernels (csrc/patchify.cu) against the oracle pinned to the reference: exact index
maps, gathers and scatters on the golden fixtures and on larger random meshes, plus the reference's own
round-trip test (utils/modular_testing.py:7-41: inverse(patchify(x)) == x) at full size."""
import numpy as np
import pytest
import torch

from oracle import patchify_oracle as po
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["uniform", "clustered"])
def test_patchify_matches_reference_goldens(cuda, tag):
    from sea_b200.patchify import DataPartitioner2D
    g = load_golden("patchify_small")
    x, y, vars_ = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_vars"]
    part = DataPartitioner2D(torch.from_numpy(x), torch.from_numpy(y), m=9, n=9, pad_id=-1, pad_field_value=0, device=cuda)
    padded, imap = part.create_partitions([torch.from_numpy(v) for v in vars_])
    assert len(padded) == 64 and len(imap) == 64
    assert np.array_equal(torch.stack(imap, 0).cpu().numpy(), g[f"{tag}_index_map"])
    assert np.array_equal(torch.stack([p[1] for p in padded], 1).cpu().numpy(), g[f"{tag}_fields"])
    assert np.array_equal(torch.stack([p[0] for p in padded], 0).cpu().numpy(), g[f"{tag}_coords"])
    rc, rf = part.inverse_partition(padded)
    assert np.array_equal(rf.cpu().numpy(), g[f"{tag}_recon"])
    assert np.array_equal(rc.cpu().numpy(), np.stack([x, y], 1))
    # the SpatialModel layout [S, P, F, C] is the same data transposed
    pfc = part.gather([torch.from_numpy(v) for v in vars_], layout_pfc=True)
    assert torch.equal(pfc, part.stacked_fields.permute(0, 1, 3, 2))
    assert torch.equal(part.scatter(pfc, layout_pfc=True), rf)


@pytest.mark.parametrize("N,S,F,m,n", [(5000, 7, 3, 9, 9), (100_003, 5, 3, 9, 9), (2048, 1, 1, 5, 13), (64, 2, 8, 9, 9)])
def test_patchify_random_meshes_vs_oracle(cuda, N, S, F, m, n):
    from sea_b200.patchify import DataPartitioner2D
    rng = np.random.RandomState(N + S)
    x = (rng.rand(N) ** 2 * 3.0 - 1.0).astype(np.float32)      # non-uniform density: ragged patches
    y = (rng.randn(N) * 0.3).astype(np.float32)
    vars_ = [rng.randn(S, N).astype(np.float32) for _ in range(F)]
    imap, counts = po.index_map(x, y, m, n, -1)
    part = DataPartitioner2D(torch.from_numpy(x), torch.from_numpy(y), m=m, n=n, pad_id=-1, pad_field_value=0, device=cuda)
    fields = part.gather([torch.from_numpy(v) for v in vars_])
    assert np.array_equal(part.index_map_tensor.cpu().numpy(), imap)
    assert np.array_equal(part.counts.cpu().numpy(), counts)
    assert np.array_equal(fields.cpu().numpy(), po.gather(vars_, imap, 0.0))
    assert np.array_equal(part.scatter(fields).cpu().numpy(), np.stack(vars_, 2))


def test_patchify_round_trip_full_size(cuda):
    """The reference's own check (unit_test_create_partitions2D) at a production-size mesh: 2048 snapshots
    of a 60k-cell mesh, 3 fields — inverse(patchify(x)) == x exactly."""
    from sea_b200.patchify import DataPartitioner2D
    g = torch.Generator(device="cuda").manual_seed(1)
    N, S, F = 60_000, 2048, 3
    x = torch.rand(N, device=cuda, generator=g) * 2.2
    y = torch.rand(N, device=cuda, generator=g) * 0.41
    vars_ = [torch.randn(S, N, device=cuda, generator=g) for _ in range(F)]
    part = DataPartitioner2D(x, y, device=cuda)
    fields = part.gather(vars_)
    assert fields.shape[:2] == (S, 64) and fields.shape[3] == F
    rec = part.scatter(fields)
    assert torch.equal(rec, torch.stack(vars_, 2))
    # padded slots hold pad_field_value, every cell appears exactly once
    im = part.index_map_tensor
    assert int((im >= 0).sum()) == N and torch.equal(torch.sort(im[im >= 0])[0], torch.arange(N, device=cuda))
    assert float(fields[:, im < 0].abs().max()) == 0.0 if bool((im < 0).any()) else True


@pytest.mark.parametrize("tag", ["uniform", "clustered"])
def test_patchify3d_matches_reference_goldens(cuda, tag):
    """DataPartitioner3D mirror (utils/data_processors.py:114-223) on the reference's golden fixtures: bit-exact."""
    from sea_b200.patchify import DataPartitioner3D
    g = load_golden("patchify3d_small")
    x, y, z, vars_ = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_z"], g[f"{tag}_vars"]
    m, n, k = (int(v) for v in g[f"{tag}_mnk"])
    part = DataPartitioner3D(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(z),
                             [torch.from_numpy(v) for v in vars_], m=m, n=n, k=k, pad_id=-1, pad_field_value=0, device=cuda)
    padded, imap = part.create_partitions()
    P = (m - 1) * (n - 1) * (k - 1)
    assert len(padded) == P and len(imap) == P
    assert np.array_equal(torch.stack(imap, 0).cpu().numpy(), g[f"{tag}_index_map"])
    assert np.array_equal(torch.stack([p[1] for p in padded], 1).cpu().numpy(), g[f"{tag}_fields"])
    assert np.array_equal(torch.stack([p[0] for p in padded], 0).cpu().numpy(), g[f"{tag}_coords"])
    rc, rf = part.inverse_partition(padded)
    assert np.array_equal(rf.cpu().numpy(), g[f"{tag}_recon"])
    assert np.array_equal(rc.cpu().numpy(), np.stack([x, y, z], 1))


@pytest.mark.parametrize("N,S,F,mnk", [(20_000, 4, 3, (9, 9, 9)), (3001, 2, 1, (3, 17, 2)), (200_000, 3, 4, (9, 9, 9))])
def test_patchify3d_random_meshes_vs_oracle(cuda, N, S, F, mnk):
    from sea_b200.patchify import DataPartitioner3D
    rng = np.random.RandomState(N)
    x = (rng.rand(N) ** 2 * 3.0 - 1.0).astype(np.float32)
    y = (rng.randn(N) * 0.3).astype(np.float32)
    z = (rng.rand(N) * 5.0).astype(np.float32)
    vars_ = [rng.randn(S, N).astype(np.float32) for _ in range(F)]
    m, n, k = mnk
    imap, counts = po.index_map3d(x, y, z, m, n, k, -1)
    part = DataPartitioner3D(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(z), [torch.from_numpy(v) for v in vars_],
                             m=m, n=n, k=k, device=cuda)
    fields = part.gather(part.var_list)
    assert np.array_equal(part.index_map_tensor.cpu().numpy(), imap)
    assert np.array_equal(part.counts.cpu().numpy(), counts)
    assert np.array_equal(fields.cpu().numpy(), po.gather(vars_, imap, 0.0))
    assert np.array_equal(part.scatter(fields).cpu().numpy(), np.stack(vars_, 2))   # the reference's own round-trip check

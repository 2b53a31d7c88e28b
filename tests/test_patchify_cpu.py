"""The numpy restatement of DataPartitioner2D against golden vectors of the unmodified reference
(oracle/make_golden_patchify.py): index maps, padded fields and the inverse, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import patchify_oracle as po
from tests.helpers import load_golden


@pytest.mark.parametrize("tag", ["uniform", "clustered"])
def test_patchify_oracle_matches_reference(tag):
    g = load_golden("patchify_small")
    x, y, vars_ = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_vars"]
    imap, counts = po.index_map(x, y, 9, 9, -1)
    assert np.array_equal(imap, g[f"{tag}_index_map"])
    assert counts.sum() == x.size and imap.shape[1] == counts.max()
    fields = po.gather(list(vars_), imap, 0.0)
    assert np.array_equal(fields, g[f"{tag}_fields"])
    rec = po.scatter(fields, imap, x.size)
    assert np.array_equal(rec, g[f"{tag}_recon"])
    if tag == "clustered":
        assert (counts == 0).sum() > 0          # the fixture has empty patches


def test_linspace_restatement_matches_torch():
    rng = np.random.RandomState(0)
    for _ in range(200):
        lo, hi = np.float32(rng.randn()), np.float32(rng.randn() + 3.0)
        for steps in (2, 5, 9, 10, 33):
            ref = torch.linspace(torch.tensor(lo), torch.tensor(hi), steps).numpy()
            assert np.array_equal(po.linspace_f32(lo, hi, steps), ref), (lo, hi, steps)


@pytest.mark.parametrize("tag", ["uniform", "clustered"])
def test_patchify3d_oracle_matches_reference(tag):
    """DataPartitioner3D (utils/data_processors.py:114-223): index map, padded fields and inverse, bit for bit."""
    g = load_golden("patchify3d_small")
    x, y, z, vars_ = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_z"], g[f"{tag}_vars"]
    m, n, k = (int(v) for v in g[f"{tag}_mnk"])
    imap, counts = po.index_map3d(x, y, z, m, n, k, -1)
    assert imap.shape[0] == (m - 1) * (n - 1) * (k - 1)
    assert np.array_equal(imap, g[f"{tag}_index_map"])
    assert counts.sum() == x.size and imap.shape[1] == counts.max()
    fields = po.gather(list(vars_), imap, 0.0)
    assert np.array_equal(fields, g[f"{tag}_fields"])
    assert np.array_equal(po.scatter(fields, imap, x.size), g[f"{tag}_recon"])

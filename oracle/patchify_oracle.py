"""CPU restatement (numpy) of the reference's DataPartitioner2D / DataPartitioner3D — TEST INFRASTRUCTURE ONLY (only tests/
may import it).  Follows utils/data_processors.py:21-59 (create_partitions), :61-88 (pad_partitions)
and :90-111 (inverse_partition).  Pinned against the unmodified reference by
oracle/make_golden_patchify.py -> tests/golden/patchify_small.npz (tests/test_patchify_cpu.py)."""
import numpy as np


def linspace_f32(lo, hi, steps):
    """torch.linspace in fp32 (ATen RangeFactories: forward from `start` in the first half, backward
    from `end` in the second, step = (end - start) / (steps - 1) in float; `start + step * i` is a fused
    multiply-add in both the CPU and the CUDA build, i.e. ONE rounding — emulated here in float64, where
    the product of two floats is exact)."""
    lo, hi = np.float32(lo), np.float32(hi)
    step = np.float32((hi - lo) / np.float32(steps - 1))
    out = np.empty(steps, dtype=np.float32)
    half = steps // 2
    for i in range(steps):
        out[i] = np.float32(np.float64(lo) + np.float64(step) * i) if i < half else \
                 np.float32(np.float64(hi) - np.float64(step) * (steps - i - 1))
    return out


def index_map(x, y, m=9, n=9, pad_id=-1):
    """-> (index_map [P, C] int64 padded with pad_id, counts [P]).  :27-59, :61-88."""
    x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
    xb = linspace_f32(x.min(), x.max(), m)
    yb = linspace_f32(y.min(), y.max(), n)
    ix = np.clip(np.searchsorted(xb, x, side="right"), 1, m - 1)     # bucketize(right=True), clamp_
    iy = np.clip(np.searchsorted(yb, y, side="right"), 1, n - 1)
    lists = []
    for i in range(1, m):
        for j in range(1, n):
            lists.append(np.nonzero((ix == i) & (iy == j))[0].astype(np.int64))
    cap = max(len(l) for l in lists)
    out = np.full((len(lists), cap), pad_id, dtype=np.int64)
    for p, l in enumerate(lists):
        out[p, : len(l)] = l
    return out, np.array([len(l) for l in lists], dtype=np.int32)


def index_map3d(x, y, z, m=9, n=9, k=9, pad_id=-1):
    """DataPartitioner3D.create_partitions + pad_partitions (utils/data_processors.py:132-196): patches ordered
    i (x) outermost, then j (y), then k (z)."""
    x, y, z = (np.asarray(a, np.float32) for a in (x, y, z))
    ix = np.clip(np.searchsorted(linspace_f32(x.min(), x.max(), m), x, side="right"), 1, m - 1)
    iy = np.clip(np.searchsorted(linspace_f32(y.min(), y.max(), n), y, side="right"), 1, n - 1)
    iz = np.clip(np.searchsorted(linspace_f32(z.min(), z.max(), k), z, side="right"), 1, k - 1)
    lists = []
    for i in range(1, m):
        for j in range(1, n):
            for l in range(1, k):
                lists.append(np.nonzero((ix == i) & (iy == j) & (iz == l))[0].astype(np.int64))
    cap = max(len(l) for l in lists)
    out = np.full((len(lists), cap), pad_id, dtype=np.int64)
    for p, l in enumerate(lists):
        out[p, : len(l)] = l
    return out, np.array([len(l) for l in lists], dtype=np.int32)


def gather(vars_, imap, pad_value=0.0):
    """vars_: list of [S, N] -> [S, P, C, F]  (:47, :73-75, stacked at :529)."""
    v = np.stack([np.asarray(a, np.float32) for a in vars_], axis=2)       # [S, N, F]
    S, N, F = v.shape
    P, C = imap.shape
    out = np.full((S, P, C, F), pad_value, dtype=np.float32)
    for p in range(P):
        valid = imap[p] >= 0
        out[:, p, valid, :] = v[:, imap[p, valid], :]
    return out


def scatter(part, imap, n_cells):
    """[S, P, C, F] -> [S, N, F]  (:101-109)."""
    S, P, C, F = part.shape
    out = np.zeros((S, n_cells, F), dtype=np.float32)
    for p in range(P):
        valid = imap[p] >= 0
        out[:, imap[p, valid], :] = part[:, p, valid, :]
    return out

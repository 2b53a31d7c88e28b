"""Host-side logic of the N>1 path on CPU: gloo, world_size 2 (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from sea_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    data = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)     # 7 trajectories over 2 ranks
    mine = parallel.shard_trajectories(data, r, w)
    flat = torch.full((10,), float(rank + 1))
    parallel.allreduce_mean_(flat)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    if rank == 0:
        q.put((flat.tolist(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_gradient_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert flat == [1.5] * 10                                   # mean of 1 and 2
    rows = gathered[0] + gathered[1]
    assert len(gathered[0]) == 4 and len(gathered[1]) == 3      # balanced contiguous shards
    assert rows == torch.arange(21, dtype=torch.float32).view(7, 3).tolist()


@pytest.mark.parametrize("n,world", [(256, 8), (7, 2), (3, 4), (0, 2)])
def test_shard_range_partitions(n, world):
    from sea_b200.parallel import shard_range
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


def test_flat_gradient_layout_follows_backward_completion_order():
    """The DP exchange sends each bucket as soon as the backward has finished it (sea_temporal_desc.bwd_events): the flat
    buffer holds the GEMM weight gradients first, grouped in completion order, and the small reduction-produced
    gradients (the only region zero-filled after zero_grad) as one contiguous tail."""
    from sea_b200 import _structs as S
    from sea_b200.temporal import TemporalModel
    m = TemporalModel(1, 64, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln")
    eng = m.engine()
    flat = eng.flat_grad()
    buckets = eng.grad_buckets()
    assert len(buckets) == S.BWD_GROUPS + 1
    assert buckets[0][0] == 0 and buckets[-1][1] == flat.numel()
    assert all(buckets[i][1] == buckets[i + 1][0] for i in range(len(buckets) - 1))

    def names(k):
        b, e = buckets[k]
        lo, hi = flat.data_ptr() + 4 * b, flat.data_ptr() + 4 * e
        return sorted(n for n, v in eng._grad_views.items() if lo <= v.data_ptr() < hi)

    assert names(0) == sorted(f"ln.{i}.cond_mlp.2.weight" for i in range(2))
    assert names(1) == sorted([f"blocks.0.mlp.{i}.layers.{j}.weight" for i in range(2) for j in (0, 3)]
                              + [f"blocks.0.proj.{i}.weight" for i in range(2)])
    assert names(2) == sorted(f"blocks.0.ln.exp.{i}.2.cond_mlp.2.weight" for i in range(2))
    assert all("cross" in n for n in names(3)) and len(names(3)) > 0
    assert all(".attn.self." in n or ".ln.exp." in n for n in names(4)) and len(names(4)) > 0
    # the tail: everything that is not a weight-gradient GEMM output
    assert all(v.dim() != 2 or "ib.layers" in n or "cond_mlp.0." in n
               for n, v in eng._grad_views.items() if n in names(5))
    assert sum(len(names(k)) for k in range(6)) == len(eng._grad_views)


def test_plan_buckets_merges_small_groups():
    from sea_b200.parallel import plan_buckets
    groups = [(0, 10), (10, 110), (110, 120), (120, 125), (125, 225)]
    assert plan_buckets(groups, 2, min_bytes=100) == [(0, 110, 1), (110, 225, 4)]
    assert plan_buckets(groups, 2, min_bytes=1) == [(b, e, k) for k, (b, e) in enumerate(groups)]
    assert plan_buckets(groups, 2, min_bytes=10 ** 9) == [(0, 225, 4)]
    assert plan_buckets([(0, 100), (100, 100), (100, 105)], 4, min_bytes=100) == [(0, 105, 2)]
    assert plan_buckets([(0, 0)], 4) == []


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from sea_b200 import parallel
    parallel.init_from_env("gloo")

    class Eng:   # host-side stand-in: CPU buffers take the non-overlapped path (no CUDA stream to overlap on)
        def __init__(self):
            self.flat = torch.full((12,), float(rank + 1))

        def flat_grad(self):
            return self.flat

    e = Eng()
    parallel.exchange_gradients(e, events=None)
    if rank == 0:
        q.put(e.flat.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_exchange_gradients_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [1.5] * 12


def test_rollout_micro_batch_ranges():
    from sea_b200.rollout import _split_ranges
    assert _split_ranges(32, 1) == [(0, 32)]
    assert _split_ranges(32, 2) == [(0, 16), (16, 32)]
    assert _split_ranges(7, 4) == [(0, 2), (2, 4), (4, 6), (6, 7)]
    assert _split_ranges(3, 8) == [(0, 1), (1, 2), (2, 3)]          # at most B groups, at most 4 (stream-K workspaces)
    assert _split_ranges(5, 0) == [(0, 5)]


def test_rollout_graph_groups():
    """Steps per CUDA graph of the rollout plans (sea_b200/rollout.py): every step exactly once, in order."""
    from sea_b200.rollout import _graph_groups
    assert _graph_groups(1, 100, 10) == [(1 + 10 * i, 10 + 10 * i) for i in range(10)]
    assert _graph_groups(1, 100, 1) == [(t, t) for t in range(1, 101)]
    assert _graph_groups(1, 100, 0) == [(1, 100)]
    assert _graph_groups(0, 6, 3) == [(0, 2), (3, 5), (6, 6)]
    for first, last, per in [(1, 7, 10), (0, 0, 4), (1, 23, 8)]:
        g = _graph_groups(first, last, per)
        assert [t for a, b in g for t in range(a, b + 1)] == list(range(first, last + 1))

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


def test_flat_gradient_layout_puts_stream_mlp_bucket_last():
    """The DP exchange overlaps the bucket whose gradients are final first in the backward (the stream-MLP
    weights, sea_temporal_backward_milestone): it must be one contiguous tail of the flat buffer, behind the
    atomically accumulated (zero-filled) head and the other GEMM weights."""
    from sea_b200.temporal import TemporalModel
    m = TemporalModel(1, 64, 2, 64, 2, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln")
    eng = m.engine()
    flat, k = eng.flat_grad(), eng.mlp_grad_offset()
    tail = sorted(n for n, v in eng._grad_views.items() if v.data_ptr() >= flat.data_ptr() + 4 * k)
    assert tail == sorted(f"blocks.0.mlp.{i}.layers.{j}.weight" for i in range(2) for j in (0, 3))
    assert eng._small_elems <= k < flat.numel()
    assert (flat.numel() - k) == 4 * 64 * 128          # 2 streams x (E x H + H x E), E=64, H=2E


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

        def mlp_grad_offset(self):
            return 8

    e = Eng()
    parallel.exchange_gradients(e, armed_event=None)
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

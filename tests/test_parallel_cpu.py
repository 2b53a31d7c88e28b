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

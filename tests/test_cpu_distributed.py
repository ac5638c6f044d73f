"""world_size-2 gloo test of the N>1 host logic (sharding by global sample index + all-gather)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = bench.shard_range(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1, 1).expand(-1, 1, 2, 2).contiguous()
    full = bench.gather_images(local, n_total, world)
    ok = torch.equal(full[:, 0, 0, 0], torch.arange(n_total, dtype=torch.float32))
    t = bench.max_over_ranks(float(rank + 1), torch.device("cpu"))
    out.put((rank, ok, t, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, 10, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=100) for _ in ps)
    [p.join(60) for p in ps]
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]          # max over ranks
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 5, 5, 10)


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import bench
    for n, w in ((65536, 8), (10, 3), (7, 8), (1024, 1)):
        spans = [bench.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))

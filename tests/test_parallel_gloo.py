"""world_size-2 gloo tests (CPU) of the sample sharding and the final gathers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpflow_b200 import parallel as PL


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 10, 100000, 12501):
        for world in (1, 2, 3, 4, 8):
            spans = [PL.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def test_round_robin_covers_all():
    for n in (0, 1, 5, 16):
        for world in (1, 2, 4):
            got = sorted(i for r in range(world) for i in PL.round_robin(n, r, world))
            assert got == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert PL.rank_world() == (rank, world)
        # ragged candidate records: rank 0 has 3 rows, rank 1 none, ...
        n_r = [3, 0][rank] if world == 2 else rank
        rows = torch.arange(n_r * 4, dtype=torch.float32).reshape(n_r, 4) + 100 * rank
        allr = PL.gather_rows(rows)
        exp = torch.cat([torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) + 100 * r
                         for r, n in enumerate([3, 0] if world == 2 else range(world))])
        assert torch.equal(allr, exp)
        # empty everywhere
        assert PL.gather_rows(torch.zeros(0, 2)).shape == (0, 2)
        # shard -> process -> gather reproduces the single-process result
        total = 11
        first, count = PL.shard_range(total)
        local = (torch.arange(first, first + count, dtype=torch.float64) ** 2)[:, None]
        full = PL.gather_rows(local)
        assert torch.equal(full[:, 0], torch.arange(total, dtype=torch.float64) ** 2)
        # round-robin deal and its inverse
        n_items = 7
        mine = PL.round_robin(n_items)
        vals = torch.tensor([[10.0 * i] for i in mine])
        back = PL.gather_round_robin(vals, n_items)
        assert torch.equal(back[:, 0], 10.0 * torch.arange(n_items))
        assert PL.all_sum(rank + 1) == world * (world + 1) / 2
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gathers_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

"""CPU test of the N>1 host logic (gloo, world_size 2): contiguous cloud sharding + the final metric all_gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(REPO, "point-sam_b200"))
    from psam_b200.parallel import gather_metric, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(total, rank, world)
    local = torch.stack([torch.tensor([float(i), float(i) ** 2]) for i in range(a, b)]) if b > a else torch.zeros(0, 2)
    full = gather_metric(local, total)
    q.put((rank, (a, b), full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [32, 7])
def test_shard_and_gather_gloo(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), float(i) ** 2] for i in range(total)]
    ranges = [r[1] for r in res]
    assert ranges[0][0] == 0 and ranges[-1][1] == total and ranges[0][1] == ranges[1][0]
    for _, _, full in res:
        assert full == want


def test_shard_range_properties():
    sys.path.insert(0, os.path.join(REPO, "point-sam_b200"))
    from psam_b200.parallel import shard_range

    for total in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1

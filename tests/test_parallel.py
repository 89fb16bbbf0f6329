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


# ---- evaluation driver sharded by crops (host logic only: a stub model stands in for the CUDA path) ----------------
class _StubGrouper:
    num_groups, group_size = 0, 0


class _StubModel(torch.nn.Module):
    """Deterministic per-cloud outputs on CPU: iteration t predicts the ground truth shifted by t points."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.prompt_iters = 3
        self.pc_encoder = type("E", (), {"patch_embed": type("P", (), {"grouper": _StubGrouper()})()})()

    def forward(self, coords, features, gt_masks, is_eval=False):
        gt = gt_masks.flatten(0, 1).float()
        return [{"prompt_masks": torch.roll(gt, t * 7, dims=1) * 2 - 1} for t in range(self.prompt_iters)]


def _write_crops(tmpdir, n):
    import numpy as np

    sys.path.insert(0, os.path.join(REPO, "point-sam_b200"))
    from pc_sam.utils import ply

    files = []
    for i in range(n):
        r = np.random.default_rng(i)
        m = 200 + 10 * i
        files.append(os.path.join(tmpdir, f"{'car' if i % 2 else 'tree'}_{i:03d}.ply"))
        ply.write_ply(files[-1], {"x": r.normal(size=m).astype(np.float32), "y": r.normal(size=m).astype(np.float32),
                                  "z": r.normal(size=m).astype(np.float32), "R": r.integers(0, 256, m).astype(np.uint8),
                                  "G": r.integers(0, 256, m).astype(np.uint8), "B": r.integers(0, 256, m).astype(np.uint8),
                                  "label": (r.random(m) < 0.4).astype(np.int32)})
    return files


def _eval_worker(rank, world, port, files, q):
    sys.path.insert(0, os.path.join(REPO, "point-sam_b200"))
    from evaluation import eval_kitti

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = eval_kitti.evaluate(_StubModel(), files, log=None)
    q.put((rank, res["total"].tolist(), {k: v.tolist() for k, v in res["per_object"].items()}, res["object_mean"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(REPO, "point-sam_b200"))
    from evaluation import eval_kitti

    files = _write_crops(str(tmp_path), 5)
    single = eval_kitti.evaluate(_StubModel(), files, log=None)
    assert single["total"].shape == (3,) and set(single["per_object"]) == {"car", "tree"}
    assert abs(single["total"][0] - 1.0) < 1e-6 and single["total"][1] < 1.0  # iteration 0 reproduces the ground truth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, files, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, total, per, omean in res:
        assert total == pytest.approx(single["total"].tolist(), abs=1e-6)
        assert omean == pytest.approx(single["object_mean"].tolist(), abs=1e-6)
        for k in per:
            assert per[k] == pytest.approx(single["per_object"][k].tolist(), abs=1e-6)


def test_c3_graph_chunk_plan_for_every_rank_count():
    """bench.py config c3: 32 clouds over 1 / 2 / 4 / 8 ranks, four graphs in flight per rank, at most 4 clouds per graph."""
    from psam_b200.parallel import plan_graph_chunks, shard_range

    for world, want in ((1, (4, 8)), (2, (4, 4)), (4, (2, 4)), (8, (1, 4))):
        for rank in range(world):
            lo, hi = shard_range(32, rank, world)
            assert plan_graph_chunks(hi - lo, 4, 4) == want, (world, rank)
    assert plan_graph_chunks(7, 4, 4) == (1, 7)      # prime shard: single-cloud graphs
    assert plan_graph_chunks(6, 4, 4) == (1, 6)
    assert plan_graph_chunks(12, 2, 4) == (4, 3)
    assert plan_graph_chunks(1, 4, 4) == (1, 1)

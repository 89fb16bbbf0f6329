"""CPU tests: the oracle against the committed golden vectors (minted from the reference's own
Python modules by oracle/make_golden.py) and against the reference test's numpy FPS oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import synth, tokenizer_ref, torch_ref
from oracle.make_golden import state_checksum


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def test_fps_fixtures(golden_dir):
    g = _load(golden_dir, "fps_cases")
    for i in range(int(g["n"])):
        B, N, G, seed = [int(v) for v in g[f"case{i}"]]
        xyz, _ = synth.make_batch(B, N, seed, str(g[f"kind{i}"]))
        want = g[f"idx{i}"].astype(np.int64)
        assert (tokenizer_ref.fps(xyz.numpy(), G) == want).all()
        assert (tokenizer_ref.fps_closed(xyz.numpy(), G) == want).all()


@pytest.mark.parametrize("shape", [(1, 31, 2), (2, 1024, 128), (3, 1025, 129), (4, 1024, 512)])
def test_fps_matches_reference_numpy_oracle(shape):
    # third_party/torkit3d/tests/ops/test_sample_farthest_points.py:54-62 (np.random.seed(0), rand)
    B, N, G = shape
    np.random.seed(0)
    pts = np.random.rand(B, N, 3)
    want = tokenizer_ref.fps_numpy(pts, G)
    assert (tokenizer_ref.fps(pts.astype(np.float32), G) == want).all()


def test_fps_tie_break_worked_example():
    # with T=32 threads, equal maxima held by "threads" 1 and 4 -> thread 4 wins
    # (bitrev5(4)=4 < bitrev5(1)=16).  Points: index 0 at origin, candidates at distance 1.
    pts = np.zeros((1, 32, 3), np.float32)
    pts[0, 1] = [1, 0, 0]
    pts[0, 4] = [0, 1, 0]
    idx = tokenizer_ref.fps(pts, 2)
    assert idx[0, 1] == 4
    assert tokenizer_ref.fps_closed(pts, 2)[0, 1] == 4
    # all points identical -> previous index repeats
    same = np.ones((1, 40, 3), np.float32)
    assert (tokenizer_ref.fps(same, 5) == 0).all()


def test_knn_matches_cdist_topk():
    xyz, _ = synth.make_batch(2, 1500, 3)
    centers = xyz[:, :40]
    idx, d2 = tokenizer_ref.knn(centers.numpy(), xyz.numpy(), 16)
    _, ref = torch_ref.knn_points(centers, xyz, 16, sorted=True)
    assert (np.sort(idx, -1) == np.sort(ref.numpy(), -1)).all()


@pytest.mark.parametrize("name", ["tiny", "tiny_fused_qkv"])
def test_model_restatement_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    B, N, G, K, P, seed = [int(v) for v in g["meta"]]
    model = torch_ref.build_model(str(g["encoder"]), G, K, seed=1234 + seed)
    assert state_checksum(model.state_dict()) == str(g["weights_checksum"]), "seeded init drifted"
    xyz, feats = torch.from_numpy(g["xyz"]), torch.from_numpy(g["feats"])
    pc, pl = torch.from_numpy(g["prompt_coords"]), torch.from_numpy(g["prompt_labels"])
    with torch.no_grad():
        emb, patches = model.pc_encoder(xyz, feats)
        masks, iou = model.predict_masks(xyz, feats, pc, pl, None, True)
        masks2, iou2 = model.predict_masks(xyz, feats, pc, pl, torch.from_numpy(g["prompt_mask"]), False)
    assert (patches["fps_idx"].numpy() == g["fps_idx"]).all()
    assert (np.sort(patches["knn_idx"].numpy(), -1) == g["knn_idx_sorted"]).all()
    np.testing.assert_allclose(emb.numpy(), g["pc_embeddings"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(masks.numpy(), g["masks"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(iou.numpy(), g["iou"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(masks2.numpy(), g["masks2"], atol=2e-5, rtol=1e-4)
    # the reference's own mm-expansion cdist stays within the north-star tolerance of the exact form
    np.testing.assert_allclose(g["masks_mm"], g["masks"], atol=1e-3, rtol=1e-2)


def test_synth_prompts_and_normalisation():
    xyz, feats = synth.make_batch(2, 512, 5)
    assert abs(float(xyz.norm(dim=-1).max()) - 1.0) < 1e-5
    assert float(feats.min()) >= -1 and float(feats.max()) <= 1
    pc, pl = synth.make_prompts(xyz, 3, 5)
    assert pc.shape == (2, 3, 3) and pl.tolist() == [[1, 0, 1], [1, 0, 1]]

"""GPU parity of the whole hot path (through the reference-shaped pc_sam API and the C ABI) against the
committed golden vectors (minted from the reference's Python modules) and the CPU oracle.

Tolerance on mask logits is the north-star bound: 1e-3 abs + 1e-2 rel (fp32); FPS indices bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth, torch_ref  # noqa: E402
from oracle.make_golden import state_checksum  # noqa: E402

ATOL, RTOL = 1e-3, 1e-2


def _build(encoder, G, K, seed):
    from pc_sam.model import build_point_sam

    oracle = torch_ref.build_model(encoder, G, K, seed=seed)
    model = build_point_sam(encoder, G, K)
    model.load_state_dict(oracle.state_dict(), strict=True)
    return model.cuda().eval(), oracle


def _report(name, got, want):
    err = (got - want).abs()
    print(f"[parity] {name}: max|err|={float(err.max()):.3e} mean|err|={float(err.mean()):.3e} "
          f"range=[{float(want.min()):.3f},{float(want.max()):.3f}]")


@pytest.mark.parametrize("name", ["tiny", "tiny_fused_qkv"])
def test_golden_end_to_end(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, N, G, K, P, seed = [int(v) for v in g["meta"]]
    model, oracle = _build(str(g["encoder"]), G, K, 1234 + seed)
    assert state_checksum(oracle.state_dict()) == str(g["weights_checksum"])
    d = torch.device("cuda:0")
    xyz, feats = torch.from_numpy(g["xyz"]).to(d), torch.from_numpy(g["feats"]).to(d)
    pc, pl = torch.from_numpy(g["prompt_coords"]).to(d), torch.from_numpy(g["prompt_labels"]).to(d)
    with torch.no_grad():
        emb, patches = model.pc_encoder(xyz, feats)
        masks, iou = model.predict_masks(xyz, feats, pc, pl, None, True)
        masks2, iou2 = model.predict_masks(xyz, feats, pc, pl, torch.from_numpy(g["prompt_mask"]).to(d), False)
    assert np.array_equal(patches["fps_idx"].cpu().numpy(), g["fps_idx"])
    assert np.array_equal(np.sort(patches["knn_idx"].cpu().numpy(), -1), g["knn_idx_sorted"])
    _report(name + " patch_embeddings", patches["embeddings"].cpu(), torch.from_numpy(g["patch_embeddings"]))
    _report(name + " pc_embeddings", emb.cpu(), torch.from_numpy(g["pc_embeddings"]))
    _report(name + " masks", masks.cpu(), torch.from_numpy(g["masks"]))
    _report(name + " masks2", masks2.cpu(), torch.from_numpy(g["masks2"]))
    np.testing.assert_allclose(patches["embeddings"].cpu().numpy(), g["patch_embeddings"], atol=2e-4, rtol=1e-3)
    np.testing.assert_allclose(emb.cpu().numpy(), g["pc_embeddings"], atol=2e-4, rtol=1e-3)
    np.testing.assert_allclose(masks.cpu().numpy(), g["masks"], atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(iou.cpu().numpy(), g["iou"], atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(masks2.cpu().numpy(), g["masks2"], atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(iou2.cpu().numpy(), g["iou2"], atol=ATOL, rtol=RTOL)
    # also within the bound of the reference's own mm-expansion cdist run
    np.testing.assert_allclose(masks.cpu().numpy(), g["masks_mm"], atol=ATOL, rtol=RTOL)


def test_golden_tie_heavy_tokenizer(golden_dir):
    """Quantised grid with duplicated points: FPS must follow the reference tie-break bit for bit."""
    g = np.load(os.path.join(golden_dir, "tiny_ties.npz"))
    B, N, G, K, P, seed = [int(v) for v in g["meta"]]
    from psam_b200 import ops

    idx, centers = ops.fps(torch.from_numpy(g["xyz"]).cuda(), G)
    assert np.array_equal(idx.cpu().numpy(), g["fps_idx"])
    assert np.array_equal(centers.cpu().numpy(), g["centers"])


def test_config1_vs_cpu_oracle():
    """BASELINE config[0]: N=4096, G=128, K=32, ViT-B, 1 prompt, plus 2 masks/cloud and the prompt loop."""
    model, oracle = _build("eva02_base_patch14_448", 128, 32, 1234)
    xyz, feats = synth.make_batch(2, 4096, 0)
    pc, pl = synth.make_prompts(xyz, 1, 0)
    d = torch.device("cuda:0")
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(xyz, feats, pc, pl, None, True)
        got_m, got_i = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d), None, True)
    _report("c1 masks", got_m.cpu(), want_m)
    np.testing.assert_allclose(got_m.cpu().numpy(), want_m.numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.numpy(), atol=ATOL, rtol=RTOL)
    # 2 masks per cloud (B*M prompt sets), 3 prompt iterations with mask feedback (config[2] semantics)
    seq_c = [synth.make_prompts(xyz, 2, s)[0].reshape(4, 1, 3) for s in (1, 2, 3)]
    seq_l = [synth.make_prompts(xyz, 2, s)[1].reshape(4, 1) for s in (1, 2, 3)]
    with torch.no_grad():
        want = oracle.predict_iterative(xyz, feats, seq_c, seq_l)
        got = model.predict_iterative(xyz.to(d), feats.to(d), [c.to(d) for c in seq_c], [l.to(d) for l in seq_l])
    for t, (a, b) in enumerate(zip(got, want)):
        _report(f"c1 iter{t} masks", a["masks"].cpu(), b["masks"])
        np.testing.assert_allclose(a["masks"].cpu().numpy(), b["masks"].numpy(), atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(a["iou_preds"].cpu().numpy(), b["iou_preds"].numpy(), atol=ATOL, rtol=RTOL)


def test_demo_api_and_errors():
    model, oracle = _build("eva02_test_tiny", 32, 16, 5)
    xyz, feats = synth.make_batch(1, 1500, 3)
    pc, pl = synth.make_prompts(xyz, 2, 3)
    d = torch.device("cuda:0")
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(xyz, feats, pc, pl, None, True)
    model.set_pointcloud(xyz.to(d), feats.to(d))
    mask, scores, logits = model.predict_masks(pc.to(d), pl.to(d), None, True)
    np.testing.assert_allclose(logits.cpu().numpy(), want_m.numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(scores.cpu().numpy(), want_i.numpy(), atol=ATOL, rtol=RTOL)
    pm = logits[0][torch.argmax(scores[0])][None, ...]
    with torch.no_grad():
        want_m2, _ = oracle.predict_masks(xyz, feats, pc, pl, pm.cpu(), False)
    mask2, scores2, logits2 = model.predict_masks(pc.to(d), pl.to(d), pm, False)
    np.testing.assert_allclose(logits2.cpu().numpy(), want_m2.numpy(), atol=ATOL, rtol=RTOL)
    # out-of-range coordinates raise ValueError like the reference (prompt_encoder.py:44-46)
    with pytest.raises(ValueError):
        model.predict_masks((xyz * 3).to(d), feats.to(d), pc.to(d), pl.to(d))
    # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        model.predict_masks(xyz, feats, pc, pl)
    # forward() with ground-truth driven prompts (eval_kitti.py:363 call form)
    gt = (xyz[..., 0] > 0.1)[:, None, :].to(d)
    model.prompt_iters = 2
    outs = model(xyz.to(d), feats.to(d), gt, is_eval=True)
    assert len(outs) == 2 and outs[0]["masks"].shape == (1, 3, 1500) and outs[1]["masks"].shape == (1, 1, 1500)
    assert outs[1]["prompt_coords"].shape == (1, 2, 3)
    # runtime mutation of the grouper (eval_kitti.py:352-362)
    model.pc_encoder.patch_embed.grouper.num_groups = 48
    model.pc_encoder.patch_embed.grouper.group_size = 8
    m3, _ = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d))
    assert m3.shape == (1, 3, 1500)


def test_config4_shapes_vs_cpu_oracle():
    """BASELINE config[3] geometry (N=131072, group_number=2048, group_size=256, KITTI-shaped cloud) with the tiny
    encoder: exercises the streaming FPS plan (cloud larger than a cluster's registers), kNN with K=256 over 131072
    keys, the unfused tensor-core attention path (2048 tokens > 512) and the 131072-point upsampling."""
    model, oracle = _build("eva02_test_tiny", 2048, 256, 77)
    xyz, feats = synth.make_batch(1, 131072, 5, "kitti")
    pc, pl = synth.make_prompts(xyz, 2, 5)
    d = torch.device("cuda:0")
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(xyz, feats, pc, pl, None, True)
        got_m, got_i = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d), None, True)
        _, patches = model.pc_encoder(xyz.to(d), feats.to(d))
    from oracle import tokenizer_ref

    assert np.array_equal(patches["fps_idx"].cpu().numpy(), tokenizer_ref.fps(xyz.numpy(), 2048))
    _report("c4-shape masks", got_m.cpu(), want_m)
    np.testing.assert_allclose(got_m.cpu().numpy(), want_m.numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.numpy(), atol=ATOL, rtol=RTOL)


def test_graph_and_pipelined_predictors_match_eager():
    """CUDA-graph replay and the multi-stream pipelined front-end return the same logits as the eager call."""
    model, _ = _build("eva02_test_tiny", 64, 16, 11)
    d = torch.device("cuda:0")
    clouds = [synth.make_batch(1, 2048, s) for s in (1, 2, 3, 4, 5)]
    prompts = [synth.make_prompts(c[0], 1, s) for s, c in enumerate(clouds)]
    want = []
    with torch.no_grad():
        for (xyz, feats), (pc, pl) in zip(clouds, prompts):
            m, i = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d), None, True)
            want.append((m.cpu(), i.cpu()))
    pp = model.make_pipelined_predictor(1, 2048, 1, depth=3)
    pp.warmup(*[t.to(d) for t in (*clouds[0], *prompts[0])])
    pp.enable_host_results(3)
    tickets = []
    for (xyz, feats), (pc, pl) in zip(clouds, prompts):
        pp.wait_lane_free(pp.count)
        if pp.count >= pp.depth:  # the host consumes the lane's previous result before the lane is reused
            t_old = pp.count - pp.depth
            m, i = pp.result(t_old, to_host=True)
            torch.testing.assert_close(m, want[t_old][0], atol=2e-5, rtol=1e-4)
        tickets.append(pp.submit(xyz.pin_memory(), feats.pin_memory(), pc.pin_memory(), pl.pin_memory(), to_host=True))
    for t in tickets[-pp.depth:]:
        m, i = pp.result(t, to_host=True)
        torch.testing.assert_close(m, want[t][0], atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(i, want[t][1], atol=2e-5, rtol=1e-4)

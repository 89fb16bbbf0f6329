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


@pytest.mark.parametrize("name,fold_ln,fold_block", [("tiny", True, True), ("tiny_fused_qkv", True, True), ("tiny", True, False),
                                                     ("tiny_fused_qkv", True, False), ("tiny", False, False)])
def test_golden_end_to_end(golden_dir, name, fold_ln, fold_block, monkeypatch):
    from psam_b200 import engine

    monkeypatch.setattr(engine, "FUSED_INNER_LN", fold_ln)  # SwiGLU.norm folded into the GEMM epilogues / separate kernel
    monkeypatch.setattr(engine, "FUSED_BLOCK_LN", fold_block)  # norm1 / norm2 / fc_norm folded (LayerNorm-free blocks) / kernels
    monkeypatch.setattr(engine, "BLOCK_LN_POLICY", "always" if fold_block else "never")
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


@pytest.mark.parametrize("tc", [True, False])
def test_decoder_patch_row_projections_tensor_core_and_simt(tc, monkeypatch):
    """The two-way transformer's projections of the G patch rows run on the tcgen05 GEMM (keys / keys + pe kept as
    split-bf16 by the LayerNorm that updates them) or on the fp32 SIMT linear: both against the oracle, 2 prompts per
    cloud x 2 masks per cloud (Z = B*M = 4) and a prompt-mask pass."""
    from psam_b200 import engine

    monkeypatch.setattr(engine, "DECODER_TC", tc)
    model, oracle = _build("eva02_test_tiny", 96, 16, 3)
    xyz, feats = synth.make_batch(2, 3000, 8)
    pc = synth.make_prompts(xyz, 4, 8)[0].reshape(4, 2, 3)
    pl = synth.make_prompts(xyz, 4, 8)[1].reshape(4, 2)
    d = torch.device("cuda:0")
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(xyz, feats, pc, pl, None, True)
        got_m, got_i = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d), None, True)
        pm = want_m[:, 1]
        want2, _ = oracle.predict_masks(xyz, feats, pc, pl, pm, False)
        got2, _ = model.predict_masks(xyz.to(d), feats.to(d), pc.to(d), pl.to(d), pm.to(d), False)
    _report(f"decoder tc={tc}", got_m.cpu(), want_m)
    np.testing.assert_allclose(got_m.cpu().numpy(), want_m.numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(got2.cpu().numpy(), want2.numpy(), atol=ATOL, rtol=RTOL)


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
    encoder: the 16-CTA register-resident FPS plan at its capacity limit (the streaming plan for larger clouds is covered
    by test_gpu_kernels.py::test_fps_streaming_plan_beyond_cluster_registers), kNN with K=256 over 131072 keys, the
    long-sequence fused attention (2048 tokens) and the 131072-point upsampling.  The full-size ViT-L variant of this
    config is test_config4_full_size_vs_fp32_oracle_on_gpu."""
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
        if pp.count >= pp.slots:  # the host consumes the slot's previous result before the slot is reused
            t_old = pp.count - pp.slots
            m, i = pp.result(t_old, to_host=True)
            torch.testing.assert_close(m, want[t_old][0], atol=2e-5, rtol=1e-4)
        tickets.append(pp.submit(xyz.pin_memory(), feats.pin_memory(), pc.pin_memory(), pl.pin_memory(), to_host=True))
    for t in tickets[-pp.slots:]:
        m, i = pp.result(t, to_host=True)
        torch.testing.assert_close(m, want[t][0], atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(i, want[t][1], atol=2e-5, rtol=1e-4)


def test_serving_predictors_raise_value_error_for_the_offending_ticket_only():
    """VERDICT r1 weak #9 / ADVICE: GraphPredictor / PipelinedPredictor must report coordinates outside [-1, 1] like the
    reference (ValueError, prompt_encoder.py:44-46) - for that ticket, not for later valid ones nor for eager calls."""
    model, _ = _build("eva02_test_tiny", 64, 16, 11)
    d = torch.device("cuda:0")
    xyz, feats = synth.make_batch(1, 2048, 1)
    pc, pl = synth.make_prompts(xyz, 1, 1)
    good = [t.to(d) for t in (xyz, feats, pc, pl)]
    bad_cloud = [(xyz * 3).to(d), feats.to(d), (pc * 3).to(d), pl.to(d)]
    bad_prompt = [xyz.to(d), feats.to(d), (pc + 5).to(d), pl.to(d)]
    pp = model.make_pipelined_predictor(1, 2048, 1, depth=2)
    pp.warmup(*good)
    t0 = pp.submit(*good)
    t1 = pp.submit(*bad_cloud)
    t2 = pp.submit(*good)      # same lane as t0
    m0, _ = pp.result(t0)
    with pytest.raises(ValueError):
        pp.result(t1)
    m2, _ = pp.result(t2)
    t3 = pp.submit(*bad_prompt)  # lane of t1: its flag was cleared on-stream
    with pytest.raises(ValueError):
        pp.result(t3)
    t4 = pp.submit(*good)
    t5 = pp.submit(*good)
    pp.result(t4), pp.result(t5)
    with torch.no_grad():
        model.predict_masks(*good)  # an unrelated eager call on the same GPU sees no stale flag
    gp = model.make_predictor(1, 2048, 1)
    gp.warmup(*good)
    gp(*bad_cloud)
    with pytest.raises(ValueError):
        gp.check()
    gp(*good)
    gp.check()
    # inputs produced on the caller's stream are ordered before the predictor's copies
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        big = torch.empty(64 << 20, device=d).normal_()  # keeps `side` busy
        late = good[0] * 1.0 + (big[:1] * 0).sum()
        with torch.no_grad():
            m_late, _ = gp(late, *good[1:])
    gp.check()
    torch.testing.assert_close(m_late, m0, atol=2e-5, rtol=1e-4)


def test_batched_prompt_sampler_vs_reference_fixture_and_oracle(golden_dir):
    """the batched border sampler against the reference's own outputs (fixture) and against the
    oracle on larger seeded inputs, including the empty-region fallbacks."""
    from pc_sam.model import prompt_sampling as ps
    from psam_b200 import ops

    d = torch.device("cuda:0")
    z = np.load(os.path.join(golden_dir, "prompt_sampler.npz"))
    xyz, gt = torch.from_numpy(z["xyz"]), torch.from_numpy(z["gt"])
    for i in range(int(z["n"])):
        pred = torch.from_numpy(z[f"pred{i}"]) if f"pred{i}" in z else None
        c, l = ps.sample_prompts_adapter(xyz.to(d), gt.to(d), pred.to(d) if pred is not None else None, is_eval=True)
        assert c.shape == (6, 1, 3) and l.shape == (6, 1) and l.dtype == torch.bool
        assert np.array_equal(c.cpu().numpy(), z[f"coords{i}"]), i
        assert np.array_equal(l.cpu().numpy(), z[f"labels{i}"]), i
    # larger seeded case vs the oracle; thresholded prediction goes through the mask input of the kernel
    B, M, N = 2, 2, 6000
    xyz, _ = synth.make_batch(B, N, 77, "kitti")
    g = torch.Generator().manual_seed(78)
    gt = torch.stack([torch.stack([(xyz[b] - xyz[b, 100 * (m + 1)]).norm(dim=-1) < 0.5 + 0.1 * m for m in range(M)]) for b in range(B)])
    pred = (gt.reshape(B * M, N).float() * 2 - 1) * (torch.rand(B * M, N, generator=g) * 2 - 0.5)
    for thr in (None, 0.6):
        want_c, want_l = torch_ref.sample_fixed_points(xyz, gt, pred, thr, False)
        c, l = ps.sample_fixed_points(xyz.to(d), gt.to(d), pred.to(d), thr, False)
        assert torch.equal(c.cpu(), want_c) and torch.equal(l.cpu(), want_l), thr
    want_c, want_l = torch_ref.sample_fixed_points(xyz, gt, pred, None, True)
    c, l = ps.sample_fixed_points(xyz.to(d), gt.to(d), pred.to(d), None, True)
    assert torch.equal(c.cpu(), want_c) and torch.equal(l.cpu(), want_l)
    # single-mask form and the batched kernel agree
    c1, l1, _ = ps.sample_furthest_points_from_border(xyz[0].to(d), gt[0, 0].to(d).long(), gt[0, 0].to(d))
    c0, l0 = ps.sample_prompts_adapter(xyz[:1].to(d), gt[:1, :1].to(d), None, is_eval=True)
    assert torch.equal(c1, c0[0]) and bool(l1[0]) == bool(l0[0, 0])
    # a mask without border (empty / full) is an error, as in the reference (torch.stack of None)
    bad = gt.clone()
    bad[1, 0] = True
    with pytest.raises(RuntimeError):
        ps.sample_prompts_adapter(xyz.to(d), bad.to(d), None, is_eval=True)
    _, _, st = ops.border_prompt(xyz.to(d), gt.to(d))
    assert int(st.item()) == 0


def test_eval_driver_and_demo_session(tmp_path):
    """the evaluation caller (binary PLY -> forward(is_eval=True) -> IoU) and the demo wire
    format (ASCII PLY, /segment JSON) on top of the CUDA path, checked against the oracle."""
    import sys

    from pc_sam.model.loss import compute_iou
    from pc_sam.utils import ply
    from psam_b200 import native as nv

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "point-sam_b200"))
    from demo.app import SegmentSession
    from evaluation import eval_kitti

    model, oracle = _build("eva02_test_tiny", 32, 16, 9)
    d = torch.device("cuda:0")
    # ---- evaluation driver ----
    files = []
    for i, n in enumerate((700, 300)):
        xyz, feats = synth.make_batch(1, n, 50 + i, "kitti")
        raw = xyz[0].numpy() * 7.5 + np.array([3.0, -2.0, 1.0], dtype=np.float32)
        rgb = ((feats[0].numpy() * 0.5 + 0.5) * 255).astype(np.uint8)
        label = (xyz[0, :, 0] > 0.05).numpy().astype(np.int32)
        f = str(tmp_path / f"car_{i:04d}.ply")
        ply.write_ply(f, {"x": raw[:, 0].copy(), "y": raw[:, 1].copy(), "z": raw[:, 2].copy(), "R": rgb[:, 0].copy(),
                          "G": rgb[:, 1].copy(), "B": rgb[:, 2].copy(), "label": label})
        files.append(f)
    model.prompt_iters = oracle.prompt_iters = 3
    res = eval_kitti.evaluate(model, files, log=None)
    assert res["total"].shape == (3,) and list(res["per_object"]) == ["car"] and np.allclose(res["object_mean"], res["total"])
    # replay the first crop: same prompts through the oracle give the same masks, hence the same IoU
    data = eval_kitti.transform_fn(eval_kitti.load_crop(files[0]), device=d)
    eval_kitti.set_group_shape(model, 700)
    outs = model(**data, is_eval=True)
    g = oracle.pc_encoder.patch_embed.grouper
    g.num_groups, g.group_size = 700, 256
    pcs = [outs[0]["prompt_coords"].cpu()] + [outs[t]["prompt_coords"][:, t:t + 1].cpu() for t in (1, 2)]
    pls = [outs[0]["prompt_labels"].cpu()] + [outs[t]["prompt_labels"][:, t:t + 1].cpu() for t in (1, 2)]
    with torch.no_grad():
        want = oracle.predict_iterative(data["coords"].cpu(), data["features"].cpu(), pcs, pls)
    for t in range(3):
        np.testing.assert_allclose(outs[t]["masks"].cpu().numpy(), want[t]["masks"].numpy(), atol=ATOL, rtol=RTOL)
    # the sampled prompts are those the reference sampler picks from the oracle's own masks
    gt = data["gt_masks"].cpu()
    c0, l0 = torch_ref.sample_prompts_eval(data["coords"].cpu(), gt, None)
    assert torch.equal(c0, pcs[0]) and torch.equal(l0, pls[0].bool())
    c1, l1 = torch_ref.sample_prompts_eval(data["coords"].cpu(), gt, want[0]["prompt_masks"])
    assert torch.equal(c1, pcs[1]) and torch.equal(l1, pls[1].bool())
    iou0 = float(compute_iou(outs[0]["prompt_masks"], data["gt_masks"].flatten(0, 1)).mean())
    assert 0.0 <= iou0 <= 1.0
    # ---- demo session ----
    model.pc_encoder.patch_embed.grouper.num_groups, model.pc_encoder.patch_embed.grouper.group_size = 32, 16
    g.num_groups, g.group_size = 32, 16
    xyz, feats = synth.make_batch(1, 600, 60, "ball")
    pts = np.concatenate([xyz[0].numpy() * 4 + 1, np.round((feats[0].numpy() * 0.5 + 0.5) * 255)], axis=1)
    scene = tmp_path / "scene.ply"
    scene.write_text("ply\nformat ascii 1.0\nelement vertex 600\nproperty float x\nproperty float y\nproperty float z\n"
                     "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" +
                     "\n".join("%f %f %f %d %d %d" % tuple(r) for r in pts) + "\n")
    sess = SegmentSession(model, device=d, output_dir=str(tmp_path / "results"))
    resp = sess.pointcloud(str(scene))
    assert len(resp["xyz"]) == 1800 and len(resp["rgb"]) == 1800
    nxyz = np.array(resp["xyz"], dtype=np.float64).reshape(-1, 3)
    assert abs(np.linalg.norm(nxyz, axis=1).max() - 1.0) < 1e-9
    cached = model._cloud
    n0 = nv.LAUNCHES[0]
    r1 = sess.segment({"prompt_point": nxyz[10].tolist(), "prompt_label": 1})
    first = nv.LAUNCHES[0] - n0
    assert len(r1["seg"]) == 600 and isinstance(r1["seg"][0], bool)
    cx, cf = torch.from_numpy(nxyz).float()[None], torch.from_numpy(np.array(resp["rgb"]).reshape(-1, 3)).float()[None]
    pp, pl = cx[:, 10:11], torch.ones(1, 1, dtype=torch.long)
    with torch.no_grad():
        m, s = oracle.predict_masks(cx, cf, pp, pl, None, True)
    b = int(torch.argmax(s[0]))
    np.testing.assert_allclose(sess.prompt_mask.cpu().numpy()[0], m[0, b].numpy(), atol=ATOL, rtol=RTOL)
    stable = m[0, b].abs() > 5e-3  # sign of logits this close to 0 is not comparable
    assert np.array_equal(np.array(r1["seg"])[stable.numpy()], (m[0, b] > 0).numpy()[stable.numpy()])
    n1 = nv.LAUNCHES[0]
    r2 = sess.segment({"prompt_point": nxyz[200].tolist(), "prompt_label": 0})
    second = nv.LAUNCHES[0] - n1
    assert len(r2["seg"]) == 600 and len(sess.prompts) == 2
    # clicks reuse the embeddings computed when the cloud was loaded (the reference re-encodes on every click)
    assert model._cloud is cached and first < 100 and second < 100, (first, second)
    assert sess.next() == {"status": "cleared"} and len(sess.masks) == 1 and sess.prompts == []
    assert sess.save() == {"status": "saved"} and os.path.exists(str(tmp_path / "results" / "scene.npy"))
    assert sess.clear() == {"status": "cleared"}


def test_iterative_graph_predictor_matches_eager_forward():
    """forward(is_eval=True) replayed as one CUDA graph (encoder + all prompt iterations, sampler included)."""
    model, oracle = _build("eva02_test_tiny", 32, 16, 11)
    d = torch.device("cuda:0")
    B, M, N = 2, 2, 1500
    model.prompt_iters = 3
    pred = model.make_iterative_predictor(B, M, N)
    clouds = []
    for s_ in (0, 1):
        xyz, feats = synth.make_batch(B, N, 90 + s_)
        gt = torch.stack([torch.stack([xyz[b, :, (m + s_) % 3] > 0.1 * m for m in range(M)]) for b in range(B)])
        clouds.append((xyz.to(d), feats.to(d), gt.to(d)))
    pred.warmup(*clouds[0])
    assert pred.graph is not None and pred.launches_per_step > 0
    for c in (clouds[1], clouds[0]):
        with torch.no_grad():
            want = model(*c, is_eval=True)
        got = pred(*c)
        assert len(got) == 3
        for t in range(3):
            assert torch.equal(got[t]["prompt_coords"], want[t]["prompt_coords"])
            assert torch.equal(got[t]["prompt_labels"], want[t]["prompt_labels"])
            assert got[t]["masks"].shape == (B * M, 3 if t == 0 else 1, N)
            torch.testing.assert_close(got[t]["masks"], want[t]["masks"], atol=2e-4, rtol=1e-4)
            torch.testing.assert_close(got[t]["prompt_masks"], want[t]["prompt_masks"], atol=2e-4, rtol=1e-4)
    # oracle replay of the last cloud with the sampled prompts
    xyz, feats, gt = clouds[0]
    pcs = [got[0]["prompt_coords"].cpu()] + [got[t]["prompt_coords"][:, t:t + 1].cpu() for t in (1, 2)]
    pls = [got[0]["prompt_labels"].cpu()] + [got[t]["prompt_labels"][:, t:t + 1].cpu() for t in (1, 2)]
    oracle.prompt_iters = 3
    with torch.no_grad():
        ow = oracle.predict_iterative(xyz.cpu(), feats.cpu(), pcs, pls)
    for t in range(3):
        np.testing.assert_allclose(got[t]["masks"].cpu().numpy(), ow[t]["masks"].numpy(), atol=ATOL, rtol=RTOL)
    # deferred validity checks still fire (after the replay)
    bad = clouds[0][2].clone()
    bad[0, 1] = True
    with pytest.raises(RuntimeError):
        pred(clouds[0][0], clouds[0][1], bad)
    with pytest.raises(ValueError):
        pred(clouds[0][0] * 4, clouds[0][1], clouds[0][2])
    pred(*clouds[0])  # flags were reset: a valid cloud passes again


@pytest.mark.parametrize("ln_policy", ["never", "always"])
def test_config2_full_size_vs_fp32_oracle_on_gpu(ln_policy, monkeypatch):
    """BASELINE config[1] at full size (N=32768, group_number=512, group_size=64, EVA02-L, 24 blocks): the CUDA path
    against the fp32 PyTorch oracle evaluated on the same GPU (cuBLAS fp32, TF32 off), same weights and inputs - with the
    LayerNorm kernels (what an eager call runs) and with the LayerNorm-free blocks (what the 8-deep pipelined predictor captures)."""
    from psam_b200 import engine

    monkeypatch.setattr(engine, "BLOCK_LN_POLICY", ln_policy)
    d = torch.device("cuda:0")
    enc, G, K, N = "eva02_large_patch14_448", 512, 64, 32768
    model, oracle = _build(enc, G, K, 21)
    xyz, feats = synth.make_batch(1, N, 33)
    pc, pl = synth.make_prompts(xyz, 1, 33)
    torch.backends.cuda.matmul.allow_tf32 = False
    oracle = oracle.to(d)
    args = [t.to(d) for t in (xyz, feats, pc, pl)]
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(*args, None, True)  # FPS through the C oracle on the host
        got_m, got_i = model.predict_masks(*args)
    _report("c2 full size", got_m, want_m)
    np.testing.assert_allclose(got_m.cpu().numpy(), want_m.cpu().numpy(), atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.cpu().numpy(), atol=ATOL, rtol=RTOL)


def _full_size_parity(enc, G, K, N, kind, seed, with_mask_pass):
    """CUDA path vs the fp32 PyTorch oracle evaluated on the same GPU (cuBLAS fp32, TF32 off), same weights and inputs;
    FPS of the oracle runs through the C restatement on the host."""
    d = torch.device("cuda:0")
    model, oracle = _build(enc, G, K, seed)
    xyz, feats = synth.make_batch(1, N, seed + 12, kind)
    pc, pl = synth.make_prompts(xyz, 1, seed + 12)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    oracle = oracle.to(d)
    args = [t.to(d) for t in (xyz, feats, pc, pl)]
    with torch.no_grad():
        want_m, want_i = oracle.predict_masks(*args, None, True)
        got_m, got_i = model.predict_masks(*args)
        _report(f"{enc} N={N} G={G} K={K} full size", got_m, want_m)
        np.testing.assert_allclose(got_m.cpu().numpy(), want_m.cpu().numpy(), atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(got_i.cpu().numpy(), want_i.cpu().numpy(), atol=ATOL, rtol=RTOL)
        if with_mask_pass:
            # second prompt iteration: the best mask of the first pass goes through the mask encoder (pc_sam.py:176-180)
            best = torch.argmax(want_i, dim=1)
            pm = want_m[torch.arange(want_m.shape[0], device=d), best]
            pc2, pl2 = synth.make_prompts(xyz, 2, seed + 13)
            want2, wi2 = oracle.predict_masks(args[0], args[1], pc2.to(d), pl2.to(d), pm, False)
            got2, gi2 = model.predict_masks(args[0], args[1], pc2.to(d), pl2.to(d), pm, False)
            _report(f"{enc} N={N} mask-encoder pass", got2, want2)
            np.testing.assert_allclose(got2.cpu().numpy(), want2.cpu().numpy(), atol=ATOL, rtol=RTOL)
            np.testing.assert_allclose(gi2.cpu().numpy(), wi2.cpu().numpy(), atol=ATOL, rtol=RTOL)


def test_config4_full_size_vs_fp32_oracle_on_gpu():
    """BASELINE config[3] at full size: EVA02-L (24 real blocks, 2048-token attention rows), N=131072, group_number=2048,
    group_size=256, KITTI-shaped cloud, plus one mask-encoder pass (524288 mini-PointNet rows per mask)."""
    _full_size_parity("eva02_large_patch14_448", 2048, 256, 131072, "kitti", 41, True)


@pytest.mark.parametrize("ln_policy", ["never", "always"])
def test_config5_full_size_vs_fp32_oracle_on_gpu(ln_policy, monkeypatch):
    """BASELINE config[4] model at full size: EVA-giant (40 blocks, 16 heads x 88, fused qkv with q/v bias, GELU MLP 6144),
    N=32768, group_number=512, group_size=64; LayerNorm kernels and LayerNorm-free blocks (GELU MLP variant of the fold)."""
    from psam_b200 import engine

    monkeypatch.setattr(engine, "BLOCK_LN_POLICY", ln_policy)
    _full_size_parity("eva_giant_patch14_560", 512, 64, 32768, "ball", 43, False)


@pytest.mark.gpu
def test_voronoi_tokenizer_and_grouper_options_vs_reference_fixture(golden_dir):
    """NNGrouper / PatchEmbedNN (Voronoi tokenizer: nearest-centre assignment, per-point residual MLPs on the tensor cores,
    maximum per cell, per-cell MLPs) and KNNGrouper(use_fps=False, centralize_features=True) against outputs of the
    reference's own modules (tests/golden/variants.npz)."""
    from oracle import torch_ref
    from oracle.make_golden import GROUPER_OPTS, VORONOI, state_checksum
    from pc_sam.model.common import KNNGrouper, NNGrouper, group_with_centers_and_nn
    from pc_sam.model.pc_encoder import PatchEmbedNN

    z = np.load(os.path.join(golden_dir, "variants.npz"))
    d = torch.device("cuda:0")
    v = VORONOI
    torch.manual_seed(4321)
    oracle = torch_ref.PatchEmbedNN(7, v["hidden"], v["out"], v["G"]).eval()
    assert state_checksum(oracle.state_dict()) == str(z["voronoi_weights_checksum"])
    m = PatchEmbedNN(7, v["hidden"], v["out"], v["G"])
    m.load_state_dict(oracle.state_dict(), strict=True)
    m = m.to(d).eval()
    xyz, feats = torch.from_numpy(z["voronoi_xyz"]).to(d), torch.from_numpy(z["voronoi_feats"]).to(d)
    with torch.no_grad():
        out = m(xyz, feats)
        grp = NNGrouper(v["G"])(xyz, feats)
        again = group_with_centers_and_nn(xyz, feats, out["centers"], out["nn_idx"])
    assert np.array_equal(out["nn_idx"].cpu().numpy(), z["voronoi_nn_idx"])
    np.testing.assert_allclose(out["centers"].cpu().numpy(), z["voronoi_centers"], atol=0)
    np.testing.assert_allclose(out["features"].cpu().numpy(), z["voronoi_features"], atol=2e-6)
    np.testing.assert_allclose(grp["features"].cpu().numpy(), z["voronoi_features"], atol=2e-6)
    np.testing.assert_allclose(again.cpu().numpy(), z["voronoi_features"], atol=2e-6)
    np.testing.assert_allclose(out["embeddings"].cpu().numpy(), z["voronoi_embeddings"], atol=2e-4, rtol=1e-4)
    # hierarchical tokenizer (PatchEmbedHier, pc_encoder.py:200-239): two KNNGrouper levels, the second on FPS-ordered centres
    from oracle.make_golden import HIER
    from pc_sam.model.pc_encoder import PatchEmbedHier

    h = HIER
    torch.manual_seed(4322)
    oh = torch_ref.PatchEmbedHier(6, h["out"], list(h["G"]), list(h["K"]), list(h["radius"])).eval()
    assert state_checksum(oh.state_dict()) == str(z["hier_weights_checksum"])
    mh = PatchEmbedHier(6, h["out"], list(h["G"]), list(h["K"]), list(h["radius"]))
    mh.load_state_dict(oh.state_dict(), strict=True)
    mh = mh.to(d).eval()
    with torch.no_grad():
        p1, p2 = mh(torch.from_numpy(z["hier_xyz"]).to(d), torch.from_numpy(z["hier_feats"]).to(d))
    np.testing.assert_allclose(p1["centers"].cpu().numpy(), z["hier_centers1"], atol=0)
    np.testing.assert_allclose(p2["centers"].cpu().numpy(), z["hier_centers2"], atol=0)
    assert np.array_equal(torch.sort(p2["knn_idx"], -1).values.cpu().numpy(), z["hier_knn2_sorted"])
    np.testing.assert_allclose(p1["embeddings"].cpu().numpy(), z["hier_emb1"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(p2["embeddings"].cpu().numpy(), z["hier_emb2"], atol=2e-4, rtol=1e-4)
    # KNNGrouper on FPS-ordered input with centralised features
    g = GROUPER_OPTS
    k = KNNGrouper(g["G"], g["K"], radius=g["radius"], centralize_features=True)
    x2, f2 = torch.from_numpy(z["grouper_xyz"]).to(d), torch.from_numpy(z["grouper_feats"]).to(d)
    with torch.no_grad():
        o2 = k(x2, f2, use_fps=False)
    assert np.array_equal(o2["fps_idx"].cpu().numpy(), z["grouper_fps_idx"])
    order = torch.argsort(o2["knn_idx"], dim=-1)
    got = torch.gather(o2["features"], 2, order.unsqueeze(-1).expand_as(o2["features"]))
    np.testing.assert_allclose(got.cpu().numpy(), z["grouper_features_sorted"], atol=2e-6)

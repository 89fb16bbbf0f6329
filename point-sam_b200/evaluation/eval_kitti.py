"""Evaluation driver with the behaviour of /root/reference/evaluation/eval_kitti.py:284-398 on top of the sm_100a
path: binary PLY crops (fields x y z R G B label) -> normalisation -> per-cloud group-count / group-size override ->
``model(**data, is_eval=True)`` (iterative GT-driven prompting) -> IoU per prompt iteration, averaged per object class
and overall.  The dataset glob and the optional rotation are arguments instead of hard-coded paths."""
from __future__ import annotations

import argparse
import glob
import os
import sys
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pc_sam.model.loss import compute_iou  # noqa: E402
from pc_sam.utils.ply import normalize_colors, normalize_points, read_ply  # noqa: E402


def transform_fn(x: Dict[str, np.ndarray], device="cuda") -> Dict[str, torch.Tensor]:
    """eval_kitti.py:91-114: one cloud with one ground-truth mask -> batched tensors."""
    xyz = normalize_points(np.asarray(x["xyz"]))
    rgb = normalize_colors(np.asarray(x["rgb"]))
    mask = np.asarray(x["mask"])
    xyz = torch.tensor(xyz, dtype=torch.float, device=device)
    rgb = torch.tensor(rgb, dtype=torch.float, device=device)
    mask = torch.tensor(mask, dtype=torch.bool, device=device)
    return {"coords": xyz[None], "features": rgb[None], "gt_masks": mask[None, None]}


def load_crop(path: str, rotation: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """eval_kitti.py:340-350: structured PLY record -> float32 xyz (optionally rotated), float32 rgb (0..255), int32 label."""
    pc = read_ply(path)
    xyz = np.column_stack([pc["x"], pc["y"], pc["z"]]).astype(np.float32)
    if rotation is not None:
        xyz = np.float32(xyz @ np.asarray(rotation, dtype=np.float64).T)
    rgb = np.column_stack([pc["R"], pc["G"], pc["B"]]).astype(np.float32)
    return {"xyz": xyz, "rgb": rgb, "mask": pc["label"].astype(np.int32)}


def set_group_shape(model, num_points: int):
    """eval_kitti.py:352-362: the tokenizer's group count / size are runtime attributes chosen per cloud."""
    g = model.pc_encoder.patch_embed.grouper
    if num_points > 30000:
        g.num_groups, g.group_size = 2048, 256
    else:
        g.num_groups, g.group_size = min(num_points, 2048), 256
        if num_points < 256:
            g.group_size = 2


def evaluate(model, files: Sequence[str], rotation: Optional[np.ndarray] = None, log=print) -> Dict[str, object]:
    """Returns {"total": [prompt_iters], "per_object": {name: [prompt_iters]}, "object_mean": [prompt_iters]}."""
    total: List[np.ndarray] = []
    per_obj: Dict[str, List[np.ndarray]] = {}
    model.eval()
    with torch.no_grad():
        for path in files:
            name = os.path.basename(path).split("_")[0]
            data = transform_fn(load_crop(path, rotation), device=next(model.parameters()).device)
            set_group_shape(model, data["coords"].shape[1])
            outputs = model(**data, is_eval=True)
            gt = data["gt_masks"].flatten(0, 1)
            ious = np.array([compute_iou(o["prompt_masks"], gt).detach().cpu().numpy().mean() for o in outputs])
            per_obj.setdefault(name, []).append(ious)
            total.append(ious)
            if log:
                log(f"Current mean IoU: {np.array(total).mean(axis=0)}")
    per = {k: np.array(v).mean(axis=0) for k, v in per_obj.items()}
    return {"total": np.array(total).mean(axis=0) if total else np.zeros(0),
            "per_object": per,
            "object_mean": np.array(list(per.values())).mean(axis=0) if per else np.zeros(0)}


def main(argv=None):
    from pc_sam.utils.checkpoint import load_model
    from pc_sam.utils.config import compose, instantiate, model_config
    from pc_sam.utils.torch_utils import replace_with_fused_layernorm

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="large")
    ap.add_argument("--config_dir", type=str, default=None, help="the reference's configs/ directory (optional)")
    ap.add_argument("--ckpt_path", type=str, default=None)
    ap.add_argument("--data", type=str, required=True, help="glob of binary PLY crops (x y z R G B label)")
    args, overrides = ap.parse_known_args(argv)
    cfg = compose(args.config_dir, args.config, overrides)["model"] if args.config_dir else model_config(args.config)
    torch.manual_seed(42)
    model = instantiate(cfg)
    model.apply(replace_with_fused_layernorm)
    if args.ckpt_path:
        load_model(model, args.ckpt_path)
    model.eval().cuda()
    res = evaluate(model, sorted(glob.glob(args.data)))
    print(f"Total mean IoU: {res['total']}")
    print(f"Object mean IoU: {res['object_mean']}")
    return res


if __name__ == "__main__":
    main()

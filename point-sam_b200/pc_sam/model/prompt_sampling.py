"""Ground-truth-driven prompt sampling for PointCloudSAM.forward(is_eval=True)
(/root/reference/pc_sam/model/common.py:287-474: sample_prompts_adapter, sample_fixed_points,
sample_furthest_points_from_border).  SURVEY.md section 8(f) rank 1 ("next" row): the control flow keeps
the reference's per-(cloud, mask) structure; the nearest-border distance (torkit3d chamfer_distance in the
reference) runs in the psam_nn_distance_f32 kernel."""
from __future__ import annotations

from typing import Union

import torch

from psam_b200 import ops


def sample_furthest_points_from_border(coords: torch.Tensor, labels: torch.Tensor, gt: torch.Tensor):
    bg_inds = labels == 0
    fg_inds = labels == 1
    if bg_inds.sum() == 0 or fg_inds.sum() == 0:
        return None, None, -1
    fg = coords[fg_inds]
    min_dists = ops.nn_distance(fg, coords[bg_inds])
    center_idx = torch.argmax(min_dists)
    return fg[center_idx][None, ...], gt[fg_inds][center_idx][None, ...], torch.max(min_dists)


@torch.no_grad()
def sample_fixed_points(points, gt_masks, pred_logits, threshold=None, from_error_region=False):
    B, M, _ = gt_masks.shape
    if pred_logits is None:
        fn = gt_masks
        fp = torch.zeros_like(fn)
    else:
        pred_logits = pred_logits.reshape(B, M, -1)
        pred_masks = pred_logits > 0 if threshold is None else pred_logits.sigmoid() > threshold
        fn = gt_masks & ~pred_masks
        fp = ~gt_masks & pred_masks
    pts, labs = [], []
    for i in range(B):
        for j in range(M):
            if from_error_region:
                c, l, _ = sample_furthest_points_from_border(points[i], (fn | fp)[i, j], gt_masks[i, j])
            else:
                pc, pl, pd = sample_furthest_points_from_border(points[i], fn[i, j], gt_masks[i, j])
                nc, nl, nd = sample_furthest_points_from_border(points[i], fp[i, j], gt_masks[i, j])
                if pd > nd:
                    c, l = pc, pl
                elif nd == -1:
                    c, l, _ = sample_furthest_points_from_border(points[i], gt_masks[i, j], gt_masks[i, j])
                else:
                    c, l = nc, nl
            pts.append(c)
            labs.append(l)
    return torch.stack(pts), torch.stack(labs)


@torch.no_grad()
def sample_prompts_adapter(points, gt_masks, pred_logits: Union[torch.Tensor, None], threshold=None, is_eval=False):
    if pred_logits is None:
        return sample_fixed_points(points, gt_masks, pred_logits, threshold, from_error_region=True)
    if not is_eval:
        raise NotImplementedError("random prompt sampling is a training-only path (out of scope)")
    return sample_fixed_points(points, gt_masks, pred_logits, threshold, from_error_region=False)

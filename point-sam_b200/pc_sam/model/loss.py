"""/root/reference/pc_sam/model/loss.py:80-98 (the evaluation metric; training losses are out of scope)."""
import torch


def compute_iou(logits: torch.Tensor, targets: torch.Tensor, threshold: float = None):
    assert logits.shape == targets.shape, (logits.shape, targets.shape)
    assert targets.dtype == torch.bool, targets.dtype
    preds = logits > 0 if threshold is None else logits.sigmoid() > threshold
    return (preds & targets).sum(-1) / (preds | targets).sum(-1)

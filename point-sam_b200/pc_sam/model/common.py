"""Tokenizer and mini-PointNet modules with the reference's import path and constructor signatures
(/root/reference/pc_sam/model/common.py:27-123, 126-187, 238-284, 477-506).  Parameters follow the
reference state-dict layout; all computation runs in the sm_100a kernels behind ``psam_b200``."""
from __future__ import annotations

import torch
from torch import nn

from psam_b200 import engine, ops


def sample_farthest_points(points: torch.Tensor, num_samples: int, transpose: bool = False) -> torch.Tensor:
    """torkit3d.ops.sample_farthest_points (ops/sample_farthest_points.py:17-31). Errors mirror the
    TORCH_CHECKs of sample_farthest_points_kernel.cu:111-115 (RuntimeError)."""
    if transpose:
        points = points.transpose(1, 2)
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor.")
    if points.dim() != 3 or points.size(2) != 3:
        raise RuntimeError("points must have shape [B, N, 3]")
    if num_samples <= 0 or points.size(1) < num_samples:
        raise RuntimeError("Check failed: 0 < num_samples <= points.size(1)")
    if points.dtype != torch.float32:
        raise RuntimeError("psam_b200 FPS computes in float32 (the reference call site passes xyz.float())")
    return ops.fps(points.contiguous(), num_samples)[0]


def batch_index_select(input, index, dim):
    """torkit3d.nn.functional.batch_index_select (nn/functional.py:34-69) - pure indexing (torch.gather)."""
    squeeze = index.dim() == 1
    if squeeze:
        index = index.unsqueeze(1)
    assert index.dim() == 2 and input.size(0) == index.size(0)
    views = [1] * input.dim()
    views[0], views[dim] = index.size(0), index.size(1)
    shape = list(input.shape)
    shape[dim] = -1
    out = torch.gather(input, dim, index.view(views).expand(shape))
    return out.squeeze(1) if squeeze else out


def fps(points: torch.Tensor, num_samples: int):
    return ops.fps(points.float().contiguous(), num_samples)[1]


def knn_points(query, key, k: int, sorted: bool = False, transpose: bool = False):
    """Same contract as the reference (returns Euclidean distances and indices); results are always sorted."""
    if transpose:
        query, key = query.transpose(1, 2), key.transpose(1, 2)
    idx, d2 = ops.knn(query.float().contiguous(), key.float().contiguous(), k, want_d2=True)
    return d2.sqrt(), idx


def compute_interp_weights(query, key, k=3, eps=1e-8):
    assert k == 3 and eps == 1e-8, "the fused kernel implements the reference defaults"
    return ops.knn3_interp(query.float().contiguous(), key.float().contiguous())


def repeat_interleave(x: torch.Tensor, repeats: int, dim: int):
    if repeats == 1:
        return x
    shape = list(x.shape)
    shape.insert(dim + 1, repeats)
    return x.unsqueeze(dim + 1).expand(shape).flatten(dim, dim + 1)


def group_with_centers_and_knn(xyz, features, centers, knn_idx, radius=None, centralize_features=False, center_idx=None):
    """common.py:126-187; centralize_features appends features[knn] - features[center_idx] (:181-185)."""
    if centralize_features and center_idx is None:
        raise RuntimeError("center_idx is required when centralize_features=True")
    return ops.group_gather(xyz.float().contiguous(), features.float().contiguous(), centers.contiguous(),
                            knn_idx.contiguous(), radius, center_idx=center_idx.contiguous() if centralize_features else None)


def group_with_centers_and_nn(xyz, features, centers, nn_idx):
    """Voronoi grouping (common.py:214-236): [unit direction to the nearest centre, distance, features] per point."""
    return ops.voronoi_features(xyz.float().contiguous(), centers.float().contiguous(), nn_idx.contiguous(),
                                features.float().contiguous())


class NNGrouper(nn.Module):
    """Group points by their nearest FPS centre (common.py:190-212)."""

    def __init__(self, num_groups: int):
        super().__init__()
        self.num_groups = num_groups

    def forward(self, xyz: torch.Tensor, features: torch.Tensor):
        return engine.run_nn_grouper(self, xyz, features)


class KNNGrouper(nn.Module):
    def __init__(self, num_groups, group_size, radius=None, centralize_features=False):
        super().__init__()
        self.num_groups = num_groups
        self.group_size = group_size
        self.radius = radius
        self.centralize_features = centralize_features

    def forward(self, xyz: torch.Tensor, features: torch.Tensor, use_fps=True):
        return engine.run_knn_grouper(self, xyz, features, use_fps)


class PatchEncoder(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_dims):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.conv1 = nn.Sequential(
            nn.Linear(in_channels, hidden_dims[0]), nn.LayerNorm(hidden_dims[0]), nn.GELU(),
            nn.Linear(hidden_dims[0], hidden_dims[0]))
        self.conv2 = nn.Sequential(
            nn.Linear(hidden_dims[0] * 2, hidden_dims[1]), nn.LayerNorm(hidden_dims[1]), nn.GELU(),
            nn.Linear(hidden_dims[1], out_channels))

    def forward(self, point_patches: torch.Tensor):
        return engine.run_patch_encoder(self, point_patches.float().contiguous())

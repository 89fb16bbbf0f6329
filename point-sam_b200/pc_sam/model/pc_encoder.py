"""/root/reference/pc_sam/model/pc_encoder.py:13-41, 84-145 (PatchEmbed, PointCloudEncoder)."""
from __future__ import annotations

import torch
from torch import nn

from psam_b200 import engine

from .common import KNNGrouper, NNGrouper, PatchEncoder


class PatchEmbed(nn.Module):
    def __init__(self, in_channels, out_channels, num_patches, patch_size, radius: float = None,
                 centralize_features=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.grouper = KNNGrouper(num_patches, patch_size, radius=radius, centralize_features=centralize_features)
        self.patch_encoder = PatchEncoder(in_channels, out_channels, [128, 512])

    def forward(self, coords: torch.Tensor, features: torch.Tensor):
        patches = self.grouper(coords, features)
        patches["embeddings"] = self.patch_encoder(patches["features"])
        return patches


class Block(nn.Module):
    """Residual MLP block of the Voronoi patch embedding (pc_encoder.py:147-162); parameters only - PatchEmbedNN runs it."""

    def __init__(self, in_channels, hidden_dim, out_channels):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_channels, hidden_dim), nn.GELU(), nn.LayerNorm(hidden_dim),
                                 nn.Linear(hidden_dim, out_channels))
        self.norm = nn.LayerNorm(out_channels)

    def forward(self, x):
        shape = x.shape
        y = x.float().reshape(-1, shape[-1]).clone()
        engine._run_res_blocks([engine._PackedResBlock(self)], y)
        return y.view(shape)


class PatchEmbedNN(nn.Module):
    """Voronoi tokenizer (pc_encoder.py:165-197): NNGrouper -> in_proj -> 3 Blocks per point -> maximum per cell ->
    3 Blocks per cell -> LayerNorm -> out_proj."""

    def __init__(self, in_channels, hidden_dim, out_channels, num_patches) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        hidden_dim = hidden_dim or out_channels
        self.grouper = NNGrouper(num_patches)
        self.in_proj = nn.Linear(in_channels, hidden_dim)
        self.blocks1 = nn.Sequential(*[Block(hidden_dim, hidden_dim, hidden_dim) for _ in range(3)])
        self.blocks2 = nn.Sequential(*[Block(hidden_dim, hidden_dim, hidden_dim) for _ in range(3)])
        self.norm = nn.LayerNorm(hidden_dim)
        self.out_proj = nn.Linear(hidden_dim, out_channels)

    def forward(self, coords: torch.Tensor, features: torch.Tensor):
        return engine.run_patch_embed_nn(self, coords, features)


class PatchEmbedHier(nn.Module):
    """PointNet++-style tokenizer with hierarchical grouping (pc_encoder.py:200-239)."""

    def __init__(self, in_channels, out_channels, num_patches, patch_size, radius=None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.grouper1 = KNNGrouper(num_patches[0], patch_size[0], radius=radius[0] if radius else None)
        self.patch_encoder1 = PatchEncoder(in_channels, 128, [64, 128])
        self.grouper2 = KNNGrouper(num_patches[1], patch_size[1], radius=radius[1] if radius else None)
        self.patch_encoder2 = PatchEncoder(128 + 3, out_channels, [128, 256])

    def forward(self, coords: torch.Tensor, features: torch.Tensor):
        return engine.run_patch_embed_hier(self, coords, features)


class PointCloudEncoder(nn.Module):
    def __init__(self, patch_embed: PatchEmbed, transformer, embed_dim: int, patch_drop_rate=0.0):
        super().__init__()
        self.transformer_dim = transformer.embed_dim
        self.embed_dim = embed_dim
        self.patch_embed = patch_embed
        self.patch_proj = nn.Linear(self.patch_embed.out_channels, self.transformer_dim)
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.transformer_dim))
        assert patch_drop_rate == 0, "PatchDropout is not compatible with decoder."
        self.patch_dropout = nn.Identity()
        self.transformer = transformer
        self.out_proj = nn.Linear(self.transformer_dim, self.embed_dim)

    def forward(self, coords, features):
        return engine.run_pc_encoder(self, coords, features)

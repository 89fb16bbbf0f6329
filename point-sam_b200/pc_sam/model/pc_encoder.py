"""/root/reference/pc_sam/model/pc_encoder.py:13-41, 84-145 (PatchEmbed, PointCloudEncoder)."""
from __future__ import annotations

import torch
from torch import nn

from psam_b200 import engine

from .common import KNNGrouper, PatchEncoder


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

"""Parameter containers for the timm EVA / EVA02 encoders Point-SAM instantiates through
``timm.create_model`` (configs/model/{base,default,giant}.yaml:10-13).  timm is an un-vendored pip
dependency that is absent offline, so this module provides ``create_model`` with the attribute and
state-dict layout of ``timm.models.eva.Eva`` (blocks.N.{norm1,attn.{q_proj,k_proj,v_proj|qkv,q_bias,
v_bias,proj},norm2,mlp.{fc1_g,fc1_x,norm,fc2|fc1,fc2}}, fc_norm, plus the unused cls_token, pos_embed,
patch_embed.proj, head).  The blocks are executed by psam_b200.engine (rope=None, no CLS/abs-pos on this
path, pc_encoder.py:136-142); if real timm is installed its modules work with the engine as well."""
from __future__ import annotations

import torch
from torch import nn

EVA_CONFIGS = {
    # name: (embed_dim, depth, heads, mlp hidden, qkv_fused, swiglu, img, patch)
    "eva02_base_patch14_448": (768, 12, 12, int(768 * 4 * 2 / 3), False, True, 448, 14),
    "eva02_large_patch14_448": (1024, 24, 16, int(1024 * 4 * 2 / 3), False, True, 448, 14),
    "eva_giant_patch14_560": (1408, 40, 16, 6144, True, False, 560, 14),
    "eva02_test_tiny": (128, 2, 4, 344, False, True, 28, 14),
    "eva_test_tiny_fused": (176, 2, 2, 256, True, False, 28, 14),
}


class EvaAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_fused):
        super().__init__()
        self.num_heads = num_heads
        if qkv_fused:
            self.qkv = nn.Linear(dim, dim * 3, bias=False)
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.register_buffer("k_bias", torch.zeros(dim), persistent=False)
            self.v_bias = nn.Parameter(torch.zeros(dim))
            self.q_proj = self.k_proj = self.v_proj = None
        else:
            self.q_proj = nn.Linear(dim, dim, bias=True)
            self.k_proj = nn.Linear(dim, dim, bias=False)
            self.v_proj = nn.Linear(dim, dim, bias=True)
            self.qkv = None
        self.proj = nn.Linear(dim, dim)


class SwiGLU(nn.Module):
    def __init__(self, dim, hidden, eps):
        super().__init__()
        self.fc1_g = nn.Linear(dim, hidden)
        self.fc1_x = nn.Linear(dim, hidden)
        self.norm = nn.LayerNorm(hidden, eps=eps)
        self.fc2 = nn.Linear(hidden, dim)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class EvaBlock(nn.Module):
    def __init__(self, dim, num_heads, hidden, qkv_fused, swiglu, eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = EvaAttention(dim, num_heads, qkv_fused)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = SwiGLU(dim, hidden, eps) if swiglu else Mlp(dim, hidden)

    def forward(self, x):
        raise RuntimeError("EvaBlock is executed by psam_b200.engine inside PointCloudEncoder.forward")


class Eva(nn.Module):
    def __init__(self, name: str):
        super().__init__()
        D, depth, heads, hidden, fused, swiglu, img, patch = EVA_CONFIGS[name]
        self.embed_dim = D
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, D, patch, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, (img // patch) ** 2 + 1, D))
        self.pos_drop = nn.Identity()
        self.blocks = nn.ModuleList([EvaBlock(D, heads, hidden, fused, swiglu) for _ in range(depth)])
        self.norm = nn.Identity()
        self.fc_norm = nn.LayerNorm(D, eps=1e-6)
        self.head = nn.Linear(D, 1000)


def create_model(model_name: str, pretrained: bool = False, **kwargs) -> Eva:
    if pretrained:
        raise RuntimeError("pretrained timm weights are not available offline")
    if model_name not in EVA_CONFIGS:
        raise RuntimeError(f"unknown model {model_name}; supported: {sorted(EVA_CONFIGS)}")
    return Eva(model_name)

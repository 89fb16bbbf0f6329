"""PointCloudSAM with the reference API (/root/reference/pc_sam/model/pc_sam.py:20-196) plus the
``set_pointcloud`` / 4-argument ``predict_masks`` form that demo/app.py:198-205 calls.

All computation runs in the sm_100a kernels behind ``psam_b200`` (no CPU / PyTorch fallback)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from psam_b200 import engine

from .common import batch_index_select, repeat_interleave
from .mask_decoder import AuxInputs, MaskDecoder
from .pc_encoder import PointCloudEncoder
from .prompt_encoder import MaskEncoder, PointEncoder


class PointCloudSAM(nn.Module):
    def __init__(self, pc_encoder: PointCloudEncoder, mask_encoder: MaskEncoder, mask_decoder: MaskDecoder,
                 prompt_iters: int, enable_mask_refinement_iterations=True):
        super().__init__()
        self.pc_encoder = pc_encoder
        self.point_encoder = PointEncoder(pc_encoder.embed_dim)
        self.mask_encoder = mask_encoder
        self.mask_decoder = mask_decoder
        self.prompt_iters = prompt_iters
        self.enable_mask_refinement_iterations = enable_mask_refinement_iterations
        self._cloud = None  # cache filled by set_pointcloud
        self._cloud_key = None

    # ------------------------------------------------------------------------------------------
    def _encode(self, coords, features):
        pc_embeddings, patches = self.pc_encoder(coords, features)
        centers = patches["centers"]
        aux = AuxInputs(coords=coords, features=features, centers=centers)
        pc_pe = engine.run_pos_embedding(self.point_encoder.pe_layer, centers, check=False)
        return dict(pc_embeddings=pc_embeddings, patches=patches, aux=aux, pc_pe=pc_pe, coords=coords)

    def _decode(self, enc, prompt_coords, prompt_labels, prompt_masks, multimask_output, center_idx=None):
        patches = enc["patches"]
        sparse = engine.run_point_encoder(self.point_encoder, prompt_coords, prompt_labels, check=False)
        dense = self.mask_encoder(prompt_masks, enc["coords"], patches["centers"], patches["knn_idx"], center_idx=center_idx)
        if prompt_masks is not None:
            dense = repeat_interleave(dense, sparse.shape[0] // dense.shape[0], 0)
        masks, iou = self.mask_decoder(enc["pc_embeddings"], enc["pc_pe"], sparse, dense, aux_inputs=enc["aux"],
                                       multimask_output=multimask_output)
        engine.raise_if_out_of_range(masks.device)  # ValueError like prompt_encoder.py:44-46 (one host sync)
        return masks, iou

    # ------------------------------------------------------------------------------------------
    def set_pointcloud(self, xyz: torch.Tensor, rgb: torch.Tensor):
        """demo/app.py:199 - encode once, keep the embeddings for subsequent prompt decodes."""
        key = (xyz.data_ptr(), xyz._version, tuple(xyz.shape), rgb.data_ptr(), rgb._version,
               self.pc_encoder.patch_embed.grouper.num_groups, self.pc_encoder.patch_embed.grouper.group_size)
        if self._cloud is not None and self._cloud_key == key:
            return  # same tensors, unmodified: keep the embeddings (the demo calls this on every click)
        with torch.no_grad():
            self._cloud = self._encode(xyz.float().contiguous(), rgb.float().contiguous())
        self._cloud_key = key
        self._cloud_keepalive = (xyz, rgb)  # the key holds data_ptr()s: keep the storages alive so they cannot be recycled

    def predict_masks(self, *args, **kwargs):
        """Two call forms:
        reference (pc_sam.py:37-45): predict_masks(coords, features, prompt_coords, prompt_labels,
            prompt_masks=None, multimask_output=True) -> (masks [B*M,C,N], iou_preds [B*M,C])
        demo (demo/app.py:200-202):  predict_masks(prompt_points [1,P,3], prompt_labels [1,P],
            prompt_mask [1,N] | None, multimask_output) -> (mask, scores, logits) after set_pointcloud()."""
        names = ("coords", "features", "prompt_coords", "prompt_labels", "prompt_masks", "multimask_output")
        demo_form = "prompt_points" in kwargs or "prompt_mask" in kwargs or (
            len(args) >= 2 and torch.is_tensor(args[1]) and args[1].dim() == 2)
        if demo_form:
            dn = ("prompt_points", "prompt_labels", "prompt_mask", "multimask_output")
            a = dict(zip(dn, args))
            a.update(kwargs)
            if self._cloud is None:
                raise RuntimeError("predict_masks(prompt_points, ...) requires set_pointcloud() first")
            with torch.no_grad():
                logits, scores = self._decode(self._cloud, a["prompt_points"], a["prompt_labels"], a.get("prompt_mask"),
                                              bool(a.get("multimask_output", True)))
            return logits, scores, logits
        a = dict(zip(names, args))
        a.update(kwargs)
        enc = self._encode(a["coords"].float().contiguous(), a["features"].float().contiguous())
        return self._decode(enc, a["prompt_coords"], a["prompt_labels"], a.get("prompt_masks"),
                            bool(a.get("multimask_output", True)))

    def make_predictor(self, batch_size: int, num_points: int, num_prompts: int, multimask_output: bool = True,
                       use_graph: bool = True):
        """Fixed-shape predictor that replays the whole path as one CUDA graph (serving entry point)."""
        from psam_b200.predictor import GraphPredictor

        return GraphPredictor(self, batch_size, num_points, num_prompts, multimask_output, use_graph)

    def make_pipelined_predictor(self, batch_size: int, num_points: int, num_prompts: int, depth: int = 3,
                                 multimask_output: bool = True, use_graph: bool = True):
        """`depth` graph predictors on separate streams, used round-robin so consecutive clouds overlap."""
        from psam_b200.predictor import PipelinedPredictor

        return PipelinedPredictor(self, batch_size, num_points, num_prompts, depth, multimask_output, use_graph)

    def make_iterative_predictor(self, batch_size: int, num_masks: int, num_points: int, use_graph: bool = True,
                                 throughput_tiles: bool = False):
        """forward(is_eval=True) - encoder, then `prompt_iters` rounds of (GT prompt sampling, prompt / mask encoders,
        decoder, best-mask feedback) - captured as ONE CUDA graph with no host synchronisation inside.
        throughput_tiles: capture with the SM-time-optimal GEMM tile policy (several predictors in flight on one GPU)."""
        from psam_b200.predictor import IterativeGraphPredictor

        return IterativeGraphPredictor(self, batch_size, num_masks, num_points, use_graph, throughput_tiles=throughput_tiles)

    # ------------------------------------------------------------------------------------------
    def predict_iterative(self, coords, features, prompt_coords_seq: List[torch.Tensor],
                          prompt_labels_seq: List[torch.Tensor]) -> List[Dict[str, torch.Tensor]]:
        """Loop body of forward (pc_sam.py:139-194) with externally supplied prompts: iteration t appends
        prompt_coords_seq[t]; multimask only at t=0; the most confident mask is fed back as prompt mask."""
        enc = self._encode(coords.float().contiguous(), features.float().contiguous())
        outs, pm = [], None
        pc, pl = prompt_coords_seq[0][:, :0], prompt_labels_seq[0][:, :0]
        for t in range(len(prompt_coords_seq)):
            pc = torch.cat([pc, prompt_coords_seq[t]], dim=1)
            pl = torch.cat([pl, prompt_labels_seq[t]], dim=1)
            masks, iou = self._decode(enc, pc, pl, pm, t == 0, center_idx=enc["patches"].get("fps_idx"))
            if t == 0:
                best = torch.argmax(iou, dim=1)
                pm = batch_index_select(masks, best, dim=1)
            else:
                best = 0
                pm = masks[:, 0]
            outs.append(dict(prompt_coords=pc, prompt_labels=pl, masks=masks, iou_preds=iou, max_iou_pred_ind=best,
                             prompt_masks=pm))
        return outs

    def forward(self, coords=None, features=None, gt_masks=None, is_eval=False, xyz=None, rgb=None, mask=None):
        """Reference forward (pc_sam.py:90-196); also accepts the xyz/rgb/mask spelling used by
        evaluation/inference.py:67-68.  Prompts are sampled from the ground truth with the reference's
        sampler (pc_sam/model/common.py:287-474) restated in pc_sam.model.prompt_sampling."""
        from .prompt_sampling import sample_prompts_adapter

        if self.training:
            # the reference's train() branch (pc_sam.py:150-165: mask-refinement iterations without new prompts, random
            # prompt sampler, autograd) is not part of this inference-only path; fail instead of silently doing eval
            raise NotImplementedError("psam_b200 is an inference-only path: call model.eval() (training mode of "
                                      "PointCloudSAM.forward is not implemented)")
        coords = coords if coords is not None else xyz
        features = features if features is not None else rgb
        gt_masks = gt_masks if gt_masks is not None else mask
        if gt_masks.dim() == 2:
            gt_masks = gt_masks.unsqueeze(1)
        gt_masks = gt_masks.bool()
        B, M = coords.shape[0], gt_masks.shape[1]
        enc = self._encode(coords.float().contiguous(), features.float().contiguous())
        outputs = []
        pc = coords.new_empty((B * M, 0, 3))
        pl = gt_masks.new_empty((B * M, 0))
        pm = None
        for i in range(self.prompt_iters):
            npc, npl = sample_prompts_adapter(coords, gt_masks, pm, is_eval=is_eval)
            pc = torch.cat([pc, npc], dim=1)
            pl = torch.cat([pl, npl], dim=1)
            masks, iou = self._decode(enc, pc, pl, pm, i == 0, center_idx=enc["patches"].get("fps_idx"))
            if i == 0:
                best = torch.argmax(iou, dim=1)
                pm = batch_index_select(masks, best, dim=1)
            else:
                best = 0
                pm = masks[:, 0]
            outputs.append(dict(prompt_coords=pc, prompt_labels=pl, masks=masks, iou_preds=iou, max_iou_pred_ind=best,
                                prompt_masks=pm))
        return outputs


PointSAM = PointCloudSAM


def build_point_sam(encoder: str = "eva02_large_patch14_448", num_patches: int = 512, patch_size: int = 64,
                    embed_dim: int = 256, prompt_iters: int = 5) -> PointCloudSAM:
    """Dependency-free mirror of configs/model/{base,default,giant}.yaml (hydra/timm are absent offline)."""
    from .eva import create_model
    from .pc_encoder import PatchEmbed
    from .transformer import TwoWayTransformer

    pe = PatchEmbed(6, 512, num_patches, patch_size)
    enc = PointCloudEncoder(pe, create_model(encoder), embed_dim)
    me = MaskEncoder(embed_dim)
    md = MaskDecoder(embed_dim, TwoWayTransformer(2, embed_dim, 8, 2048))
    return PointCloudSAM(enc, me, md, prompt_iters).eval()

"""CUDA-graph predictor: the whole hot path (tokenizer -> ViT encoder -> prompt decoder) captured once for
fixed shapes and replayed per cloud.  Inputs may be host (pinned) or device tensors; the H2D copies are
enqueued on the same stream ahead of the replay.  Public API: ``PointCloudSAM.make_predictor``."""
from __future__ import annotations

import torch

from . import engine, native as nv


class GraphPredictor:
    def __init__(self, model, B: int, N: int, P: int, multimask_output: bool = True, use_graph: bool = True,
                 device=None):
        self.model = model
        self.dev = device or next(model.parameters()).device
        self.multimask = multimask_output
        d = self.dev
        self.xyz = torch.zeros((B, N, 3), dtype=torch.float32, device=d)
        self.feats = torch.zeros((B, N, 3), dtype=torch.float32, device=d)
        self.pc = torch.zeros((B, P, 3), dtype=torch.float32, device=d)
        self.pl = torch.zeros((B, P), dtype=torch.int64, device=d)
        self.graph = None
        self.use_graph = use_graph
        self.masks = self.iou = None
        self.launches_per_step = 0
        self.stream = torch.cuda.Stream(device=d)

    def _run(self):
        enc = self.model._encode(self.xyz, self.feats)
        sparse = engine.run_point_encoder(self.model.point_encoder, self.pc, self.pl, check=False)
        dense = self.model.mask_encoder(None, self.xyz, enc["patches"]["centers"], enc["patches"]["knn_idx"])
        return self.model.mask_decoder(enc["pc_embeddings"], enc["pc_pe"], sparse, dense, aux_inputs=enc["aux"],
                                       multimask_output=self.multimask)

    def warmup(self, xyz, feats, pc, pl):
        """Eager passes (pack weights, size the allocator) then capture."""
        with torch.no_grad(), torch.cuda.stream(self.stream):
            self._load(xyz, feats, pc, pl)
            for _ in range(2):
                n0 = nv.LAUNCHES[0]
                self.masks, self.iou = self._run()
                self.launches_per_step = nv.LAUNCHES[0] - n0
            self.stream.synchronize()
            engine.raise_if_out_of_range(self.dev)
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    self.masks, self.iou = self._run()
        self.stream.synchronize()

    def _load(self, xyz, feats, pc, pl):
        self.xyz.copy_(xyz, non_blocking=True)
        self.feats.copy_(feats, non_blocking=True)
        self.pc.copy_(pc, non_blocking=True)
        self.pl.copy_(pl, non_blocking=True)

    def __call__(self, xyz, feats, pc, pl):
        """Enqueue one step on the predictor's stream; returns device tensors (valid after stream sync)."""
        with torch.no_grad(), torch.cuda.stream(self.stream):
            self._load(xyz, feats, pc, pl)
            if self.graph is not None:
                self.graph.replay()
            else:
                self.masks, self.iou = self._run()
        return self.masks, self.iou

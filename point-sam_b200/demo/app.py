"""Interactive-segmentation backend with the JSON wire format of /root/reference/demo/app.py:91-206, served by the
standard library (flask is not required).  The state machine lives in ``SegmentSession`` so it can be driven without
HTTP.  Two differences from the reference, both on the serving side of the hot path:

* the point-cloud encoder runs once per cloud - ``PointCloudSAM.set_pointcloud`` keeps the embeddings, so a click costs
  one prompt-encoder + decoder pass (the reference re-encodes the cloud on every click, app.py:199);
* the request handler never builds Python lists of per-point floats on the hot path: masks leave as one ``tolist()``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict, List, Optional

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pc_sam.utils.ply import load_ply  # noqa: E402


class SegmentSession:
    def __init__(self, model, device="cuda", output_dir="results"):
        self.sam = model
        self.device = torch.device(device)
        self.output_dir = output_dir
        self.pc_xyz: Optional[torch.Tensor] = None
        self.pc_rgb: Optional[torch.Tensor] = None
        self.obj_path: Optional[str] = None
        self.masks: List[np.ndarray] = []
        self.segment_mask: Optional[torch.Tensor] = None
        self._reset_prompts()

    def _reset_prompts(self):
        self.prompts, self.labels, self.prompt_mask = [], [], None

    def _set_cloud(self, xyz: np.ndarray, rgb: np.ndarray):
        self.pc_xyz = torch.from_numpy(np.ascontiguousarray(xyz)).to(self.device).float().unsqueeze(0)
        self.pc_rgb = torch.from_numpy(np.ascontiguousarray(rgb)).to(self.device).float().unsqueeze(0)
        self.sam.set_pointcloud(self.pc_xyz, self.pc_rgb)  # encode once; clicks reuse the embeddings
        self._reset_prompts()

    # ---- routes (names and payloads as in the reference) -------------------------------------------
    def pointcloud(self, path: str) -> Dict:
        """GET /pointcloud/<path>: load an ASCII PLY, normalise to the unit sphere, colours to 0..1 (app.py:110-140)."""
        self.obj_path = os.path.basename(path)
        pts = load_ply(path)
        xyz, rgb = pts[:, :3], pts[:, 3:6] / 255
        shift = xyz.mean(0)
        scale = np.linalg.norm(xyz - shift, axis=-1).max()
        xyz = (xyz - shift) / scale
        self._set_cloud(xyz, rgb)
        return {"xyz": xyz.flatten().tolist(), "rgb": rgb.flatten().tolist()}

    def sampled_pointcloud(self, req: Dict) -> Dict:
        """POST /sampled_pointcloud {points: {i: v}, colors: {i: v}} (app.py:91-107)."""
        pts = np.array(list(req["points"].values()), dtype=np.float64).reshape(-1, 3)
        col = np.array(list(req["colors"].values()), dtype=np.float64).reshape(-1, 3)
        self._set_cloud(pts, col)
        return {"response": "success"}

    def segment(self, req: Dict) -> Dict:
        """POST /segment {prompt_point: [x,y,z], prompt_label: 0|1} -> {seg: [bool]*N} (app.py:177-206)."""
        if self.pc_xyz is None:
            raise RuntimeError("no point cloud loaded")
        self.prompts.append(req["prompt_point"])
        self.labels.append(req["prompt_label"])
        pp = torch.tensor(self.prompts, dtype=torch.float32, device=self.device)[None]
        pl = torch.tensor(self.labels, device=self.device)[None]
        with torch.no_grad():
            self.sam.set_pointcloud(self.pc_xyz, self.pc_rgb)  # no-op for an unchanged cloud
            mask, scores, logits = self.sam.predict_masks(pp, pl, self.prompt_mask, self.prompt_mask is None)
        best = torch.argmax(scores[0])
        self.prompt_mask = logits[0][best][None]
        self.segment_mask = mask[0][best] > 0
        return {"seg": self.segment_mask.cpu().numpy().tolist()}

    def clear(self) -> Dict:
        self._reset_prompts()
        self.segment_mask = None
        return {"status": "cleared"}

    def next(self) -> Dict:
        if self.segment_mask is not None:
            self.masks.append(self.segment_mask.cpu().numpy())
        self._reset_prompts()
        return {"status": "cleared"}

    def save(self) -> Dict:
        os.makedirs(self.output_dir, exist_ok=True)
        name = (self.obj_path or "pointcloud").split(".")[0]
        np.save(os.path.join(self.output_dir, name + ".npy"),
                {"xyz": self.pc_xyz[0].cpu().numpy(), "rgb": self.pc_rgb[0].cpu().numpy(),
                 "mask": np.stack(self.masks) if self.masks else np.zeros((0, self.pc_xyz.shape[1]), dtype=bool)})
        self.masks = []
        self._reset_prompts()
        self.segment_mask = None
        return {"status": "saved"}


def make_handler(session: SegmentSession, static_dir: str, model_dir: str, pointcloud: str = None):
    """pointcloud: the file every /pointcloud/<anything> request serves (the reference always loads args.pointcloud,
    demo/app.py:91-126); None serves the requested basename from `model_dir`."""
    import threading
    from urllib.parse import urlparse

    lock = threading.Lock()  # one model, one mutable session: requests are serialised like the reference's Flask dev server
    static_root = os.path.realpath(static_dir)

    class Handler(BaseHTTPRequestHandler):
        def _json(self, obj, code=200):
            body = json.dumps(obj).encode()
            self.send_response(code)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_GET(self):
            path = urlparse(self.path).path
            if path.startswith("/pointcloud/"):
                name = pointcloud if pointcloud else os.path.basename(path)
                with lock:
                    return self._json(session.pointcloud(name if os.path.isabs(name) else os.path.join(model_dir, name)))
            rel = "index.html" if path == "/" else path.lstrip("/").replace("static/", "", 1)
            p = os.path.realpath(os.path.join(static_root, rel))
            if os.path.commonpath([p, static_root]) != static_root or not os.path.isfile(p):
                return self._json({"error": "not found"}, 404)
            with open(p, "rb") as f:
                body = f.read()
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def do_POST(self):
            n = int(self.headers.get("Content-Length") or 0)
            req = json.loads(self.rfile.read(n) or b"{}")
            routes = {"/segment": lambda: session.segment(req), "/sampled_pointcloud": lambda: session.sampled_pointcloud(req),
                      "/clear": session.clear, "/next": session.next, "/save": session.save}
            fn = routes.get(urlparse(self.path).path)
            if fn is None:
                return self._json({"error": "not found"}, 404)
            try:
                with lock:
                    out = fn()
                self._json(out)
            except (ValueError, RuntimeError) as e:
                self._json({"error": str(e)}, 400)

    return Handler


def main(argv=None):
    from pc_sam.utils.checkpoint import load_model
    from pc_sam.utils.config import compose, instantiate, model_config

    ap = argparse.ArgumentParser()
    ap.add_argument("--host", type=str, default="127.0.0.1")
    ap.add_argument("--port", type=int, default=5000)
    ap.add_argument("--pointcloud", type=str, default="scene.ply")
    ap.add_argument("--config", type=str, default="large")
    ap.add_argument("--config_dir", type=str, default=None)
    ap.add_argument("--ckpt_path", type=str, default=None)
    ap.add_argument("--static", type=str, default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "static"))
    args = ap.parse_args(argv)
    cfg = compose(args.config_dir, args.config)["model"] if args.config_dir else model_config(args.config)
    model = instantiate(cfg)
    if args.ckpt_path:
        load_model(model, args.ckpt_path)
    model.eval().cuda()
    session = SegmentSession(model)
    srv = ThreadingHTTPServer((args.host, args.port), make_handler(session, args.static, os.path.join(args.static, "models"), args.pointcloud))
    srv.serve_forever()


if __name__ == "__main__":
    main()

"""FPS + kNN sweep (BASELINE config 5): N in {8k, 32k, 128k, 512k}, G=512, K=64, B in {1, 16}; reports time and the
streaming-model / reference-equivalent GB/s against the measured HBM peak."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import synth  # noqa: E402
from psam_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
try:
    HBM = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
G, K = 512, 64


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"HBM peak {HBM} GB/s (measured copy bandwidth)")
print("| B | N | FPS ms | us/iter | FPS stream-model GB/s | % HBM | kNN ms | kNN ref-equiv GB/s | % HBM |")
print("|---|---|---|---|---|---|---|---|---|")
for B in (1, 16):
    for N in (8192, 32768, 131072, 524288):
        g = torch.Generator().manual_seed(N + B)
        xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        t_fps = timeit(lambda: ops.fps(xyz, G), 3)
        _, centers = ops.fps(xyz, G)
        t_knn = timeit(lambda: ops.knn(centers, xyz, K), 3)
        fb = (G - 1) * N * 20.0 * B
        kb = (2.0 * G * N * 4 + N * 12 + G * K * 12) * B
        print(f"| {B} | {N} | {t_fps:.3f} | {t_fps * 1e3 / (G - 1):.2f} | {fb / t_fps / 1e6:.0f} | {100 * fb / t_fps / 1e6 / HBM:.1f} | "
              f"{t_knn:.3f} | {kb / t_knn / 1e6:.0f} | {100 * kb / t_knn / 1e6 / HBM:.1f} |", flush=True)

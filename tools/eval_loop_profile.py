"""Kernel-time breakdown of the evaluation loop (forward(is_eval=True)) with torch.profiler: top kernels by device time."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from psam_b200 import synth  # noqa: E402
from pc_sam.model import build_point_sam  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
G = int(sys.argv[4]) if len(sys.argv) > 4 else 512
K = int(sys.argv[5]) if len(sys.argv) > 5 else 64
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = build_point_sam("eva02_large_patch14_448", G, K).to(dev).eval()
model.prompt_iters = iters
xyz, feats = synth.make_batch(B, N, 300, "ball")
gt = torch.stack([torch.stack([(xyz[b] - xyz[b, 997]).norm(dim=-1) < 0.45]) for b in range(B)])
args = tuple(t.to(dev) for t in (xyz, feats, gt))
with torch.no_grad():
    for _ in range(2):
        model(*args, is_eval=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(*args, is_eval=True)
        torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type is not None]
rows.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows if "psam" in e.key or "at::" in e.key or "void" in e.key)
print(f"iters={iters} B={B} N={N} G={G} K={K}")
for e in rows[:28]:
    print(f"{e.device_time_total / 1e3:9.3f} ms  n={e.count:4d}  {e.key[:110]}")

"""One eager pass of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from bench import CONFIGS  # noqa: E402
from psam_b200 import synth  # noqa: E402
from pc_sam.model import build_point_sam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--throughput-tiles", action="store_true", help="tile policy of the pipelined predictor (BN=256)")
a = ap.parse_args()
enc, N, G, K, bpg, P, kind = CONFIGS[a.config]
if a.throughput_tiles:
    from psam_b200 import ops

    ops.GEMM_TILE_HINT = 1
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = build_point_sam(enc, G, K).to(dev).eval()
xyz, feats = synth.make_batch(bpg, N, 0, kind)
pc, pl = synth.make_prompts(xyz, P, 0)
pred = model.make_predictor(bpg, N, P, True, use_graph=False)
args = [t.to(dev) for t in (xyz, feats, pc, pl)]
from psam_b200 import engine  # noqa: E402

with engine.block_ln_fold(a.throughput_tiles):  # the 8-deep pipelined predictor captures the LayerNorm-free blocks
    pred.warmup(*args)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(a.passes):
        pred(*args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step", pred.launches_per_step)

import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch
from psam_b200 import ops
dev = torch.device("cuda:0")
M, N, K = [int(v) for v in sys.argv[1:4]]
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = ops.Split(M, K, dev); a.t.normal_()
w = ops.Split(N, K, dev); w.t.normal_()
out = torch.zeros(M, N, device=dev)
for _ in range(4):
    ops.gemm(a, w, out_f32=out, passes=passes)
torch.cuda.synchronize()

"""Launch one hot kernel in isolation between cudaProfilerStart/Stop (for `ncu --profile-from-start off --set full`).
usage: python tools/kernel_once.py {attention_long|attention|knn|knn_c4|border|gemm_qkv|gemm_qkv_single|gemm_fc1|gemm_m2048|gemm_pe|gemm_rowln|fps}"""
import os
import sys
from ctypes import byref

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import synth  # noqa: E402
from psam_b200 import native as nv, ops  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1]


def attention(L, H=16, B=1, dh=64):
    D = H * dh
    qkv = ops.Split(B * L, 3 * D, dev)
    qkv.t.normal_()
    att = ops.Split(B * L, D, dev)
    mk = lambda col: qkv.operand(rows=L, k=dh, col=col, nb1=H, b1_stride=dh, nb2=B, b2_stride=L * qkv.pitch)
    qa, ka, va = mk(0), mk(D), mk(2 * D)
    return lambda: nv.check(nv.lib().psam_attention_bf16x3(byref(qa), byref(ka), byref(va), att.ptr(), att.plane, att.pitch, dh,
                                                           L * att.pitch, dh ** -0.5, nv.stream()), "attention")


def knn(N, G, K, kind):
    xyz, _ = synth.make_batch(1, N, 5, kind)
    x = xyz.to(dev)
    idx, centers = ops.fps(x, G)
    return lambda: ops.knn(centers, x, K)


def border(N):
    xyz, _ = synth.make_batch(1, N, 6, "ball")
    gt = ((xyz[0] - xyz[0, 997]).norm(dim=-1) < 0.45)[None, None]
    g = torch.Generator().manual_seed(7)
    pred = (gt.reshape(1, N).float() * 2 - 1) * (torch.rand(1, N, generator=g) * 2 - 0.3)
    x, gt, pred = xyz.to(dev), gt.to(dev), pred.to(dev)
    return lambda: ops.border_prompt(x, gt, pred, None, False)


def gemm(M, N, K, hint, variant=0):
    a, w = ops.Split(M, K, dev), ops.Split(N, K, dev)
    a.t.normal_()
    w.t.normal_()
    out = torch.zeros(M, N, device=dev)
    ops.GEMM_TILE_HINT = hint
    ops.GEMM_VARIANT = variant
    return lambda: ops.gemm(a, w, out_f32=out, passes=3)


def gemm_rowln(M=32768, N=512, K=128, group_rows=64):
    a, w = ops.Split(M, K, dev), ops.Split(N, K, dev)
    a.t.normal_()
    w.t.normal_()
    gb = torch.randn(M // group_rows, N, device=dev)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    out = ops.Split(M, N, dev)
    return lambda: ops.gemm_rowln(a, w, gamma, beta, 1e-5, out, gbias=gb, group_rows=group_rows, act=1)


def fps(N, G=512):
    xyz, _ = synth.make_batch(1, N, 5, "ball")
    x = xyz.to(dev)
    return lambda: ops.fps(x, G)


fn = {"attention_long": lambda: attention(2048), "attention": lambda: attention(512), "knn": lambda: knn(32768, 512, 64, "ball"),
      "knn_c4": lambda: knn(131072, 2048, 256, "kitti"), "border": lambda: border(32768), "gemm_qkv": lambda: gemm(512, 3072, 1024, 1),
      "gemm_pe": lambda: gemm(32768, 512, 128, 1),
      "gemm_qkv_single": lambda: gemm(512, 3072, 1024, 1, ops.GV_NO_DUAL), "gemm_fc1": lambda: gemm(512, 5504, 1024, 1),
      "gemm_m2048": lambda: gemm(2048, 3072, 1024, 1), "gemm_rowln": lambda: gemm_rowln(), "fps": lambda: fps(32768)}[what]()
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

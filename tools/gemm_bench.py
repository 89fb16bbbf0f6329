"""Micro-benchmark of psam_gemm_bf16x3 on the ViT-L shapes (tile width / split-K sweep, cold weights)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = {"qkv": (512, 3072, 1024), "proj": (512, 1024, 1024), "fc1": (512, 5504, 1024), "fc2": (512, 1024, 2752),
          "pe_conv2b": (32768, 512, 128), "pe_conv3": (32768, 512, 512), "up3": (32768, 256, 256),
          "qkv_b4": (2048, 3072, 1024), "fc1_b4": (2048, 5504, 1024), "fc2_b4": (2048, 1024, 2752)}
NW = 6


def bench(name, M, N, K, bn, sk, passes=3, iters=20):
    a = ops.Split(M, K, dev)
    a.t.normal_()
    ws = []
    for _ in range(NW):
        w = ops.Split(N, K, dev)
        w.t.normal_()
        ws.append(w)
    out = torch.zeros(M, N, device=dev)
    if bn:
        os.environ["PSAM_GEMM_BN"] = str(bn)
    else:
        os.environ.pop("PSAM_GEMM_BN", None)
    kw = dict(out_f32=out, passes=passes)
    if sk > 1:
        kw.update(accumulate=True, split_k=sk)
    for i in range(3):
        ops.gemm(a, ws[i % NW], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.gemm(a, ws[i % NW], **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    tf = 2.0 * M * N * K * passes / us / 1e6
    print(f"{name:10s} M={M:5d} N={N:5d} K={K:5d} bn={bn or 'auto':>4} split={sk} passes={passes}: {us:8.1f} us  {tf:7.1f} TFLOP/s executed", flush=True)


def bench_attn(bn):
    """S = Q K^T and O = P V^T batched over 16 heads (L=512, dh=64), as issued by the ViT block."""
    from psam_b200 import native as nv
    B, H, L, dh = 1, 16, 512, 64
    D = H * dh
    qkv = ops.Split(B * L, 3 * D, dev); qkv.t.normal_()
    s = torch.empty(B * H * L, L, device=dev)
    p = ops.Split(B * H * L, L, dev); p.t.normal_()
    vt = ops.Split(B * H * dh, L, dev); vt.t.normal_()
    att = ops.Split(B * L, D, dev)
    if bn:
        os.environ["PSAM_GEMM_BN"] = str(bn)
    else:
        os.environ.pop("PSAM_GEMM_BN", None)
    def run_s():
        qa = qkv.operand(rows=L, k=dh, col=0, nb1=H, b1_stride=dh, nb2=B, b2_stride=L * qkv.pitch)
        ka = qkv.operand(rows=L, k=dh, col=D, nb1=H, b1_stride=dh, nb2=B, b2_stride=L * qkv.pitch)
        o = ops.GemmOut(); o.out_f32, o.ldo, o.out_b1, o.out_b2, o.alpha = nv.ptr(s), L, L * L, H * L * L, 1.0
        ops.gemm_raw(qa, ka, o, 3, 1)
    def run_pv():
        pa = p.operand(rows=L, k=L, nb1=H, b1_stride=L * p.pitch, nb2=B, b2_stride=H * L * p.pitch)
        va = vt.operand(rows=dh, k=L, nb1=H, b1_stride=dh * vt.pitch, nb2=B, b2_stride=H * dh * vt.pitch)
        o = ops.GemmOut(); o.out_hi, o.out_plane, o.ldo_s, o.outs_b1, o.outs_b2, o.alpha = att.ptr(), att.plane, att.pitch, dh, L * att.pitch, 1.0
        ops.gemm_raw(pa, va, o, 3, 1)
    for name, fn in (("S", run_s), ("PV", run_pv)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:10s} batched 16 heads L=512 dh=64 bn={bn or 'auto':>4}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    if "attn" in sys.argv:
        for bn in (0, 32, 64, 128, 256):
            bench_attn(bn)
        sys.argv.remove("attn")
        if len(sys.argv) == 1:
            sys.exit(0)
    which = sys.argv[1:] or list(SHAPES)
    for name in which:
        M, N, K = SHAPES[name]
        sks = [1, 2, 4] if M <= 512 and N <= 1024 else [1]
        for sk in sks:
            for bn in (0, 64, 96, 128, 160, 192, 256):
                bench(name, M, N, K, bn, sk)
        bench(name, M, N, K, 0, 1, passes=1)

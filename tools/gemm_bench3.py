"""Graph-timed micro-benchmark of psam_gemm_bf16x3 kernel variants on the ViT-L block shapes, for one cloud (M=512) and for
coalesced requests (M=2048 / 4096), cold weights (rotating buffers), 1..8 concurrent streams.
usage: python tools/gemm_bench3.py [quick]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
NW = 6


def bench(name, M, N, K, variant, hint, sk=1, concurrent=1, swiglu=False, reps=3):
    a = ops.Split(M, K, dev)
    a.t.normal_()
    ws = [ops.Split(N, K, dev) for _ in range(NW)]
    for w in ws:
        w.t.normal_()
    outs = [torch.zeros(M, N // 2 if swiglu else N, device=dev) for _ in range(concurrent)]

    def run(i, out):
        o = ops.GemmOut()
        o.out_f32, o.ldo, o.alpha, o.tile_hint, o.variant = out.data_ptr(), (N // 2 if swiglu else N), 1.0, hint, variant
        o.swiglu = int(swiglu)
        if sk > 1:
            o.accumulate = 1
        ops.gemm_raw(a.operand(), ws[i % NW].operand(), o, 3, sk)

    streams = [torch.cuda.Stream() for _ in range(concurrent)]
    graphs = []
    for c in range(concurrent):
        with torch.cuda.stream(streams[c]):
            run(0, outs[c])
        streams[c].synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[c]):
            for i in range(12):
                run(i + c, outs[c])
        graphs.append(g)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for c in range(concurrent):
            streams[c].wait_event(e0)
            with torch.cuda.stream(streams[c]):
                for _ in range(reps):
                    graphs[c].replay()
            ev = torch.cuda.Event()
            ev.record(streams[c])
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (12 * reps * concurrent) * 1e3)
    tf = 2.0 * M * N * K * 3 / best / 1e6
    print(f"{name:14s} M={M:5d} N={N:5d} K={K:5d} var={variant:#07x} hint={hint:3d} split={sk} streams={concurrent}: {best:8.2f} us/gemm "
          f"{tf:7.1f} TF/s executed ({tf / 1460.3:.2f} of sustained peak)", flush=True)
    return tf


VARIANTS = [("oneshot", ops.GV_NO_DUAL | ops.GV_NO_PERSIST), ("oneshot2i", ops.GV_NO_DUAL | ops.GV_NO_PERSIST | ops.GV_TWO_ISSUERS), ("dual", ops.GV_DUAL | ops.GV_NO_PERSIST), ("2cta", ops.GV_2CTA | ops.GV_NO_PERSIST),
            ("persist", ops.GV_PERSIST)]
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
pe_only = len(sys.argv) > 1 and sys.argv[1] == "pe"
for M, conc in [] if pe_only else ([(4096, 1), (512, 8)] if quick else [(512, 1), (512, 8), (2048, 2), (4096, 1), (4096, 2)]):
    for nm, (N, K, sw) in {"qkv": (3072, 1024, False), "proj": (1024, 1024, False), "fc1sw": (5504, 1024, True), "fc2": (1024, 2752, False)}.items():
        for vn, var in VARIANTS:
            try:
                bench(f"{nm}/{vn}", M, N, K, var, 1, 1, conc, sw)
            except Exception as e:
                print(nm, vn, "FAILED", repr(e)[:120])
    print()
# mini-PointNet shapes (32768 rows)
for nm, (N, K) in {"pe_conv1b": (128, 128), "pe_conv2a": (512, 128), "pe_conv2b": (512, 512)}.items():
    for vn, var in VARIANTS:
        try:
            bench(f"{nm}/{vn}", 32768, N, K, var, 1, 1, 1, False)
        except Exception as e:
            print(nm, vn, "FAILED", repr(e)[:120])

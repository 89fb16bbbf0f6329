"""Graph-timed micro-benchmark of psam_gemm_bf16x3 (1-CTA vs 2-CTA, tile widths), cold weights (rotating buffers)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = {"qkv": (512, 3072, 1024), "proj": (512, 1024, 1024), "fc1": (512, 5504, 1024), "fc2": (512, 1024, 2752),
          "qkv_b4": (2048, 3072, 1024), "fc1_b4": (2048, 5504, 1024)}
NW = 8


def bench(name, M, N, K, two, bn, sk, concurrent=1, swiglu=False):
    os.environ["PSAM_GEMM_2CTA"] = str(two)
    a = ops.Split(M, K, dev)
    a.t.normal_()
    ws = [ops.Split(N, K, dev) for _ in range(NW)]
    for w in ws:
        w.t.normal_()
    outs = [torch.zeros(M, N // 2 if swiglu else N, device=dev) for _ in range(concurrent)]

    def run(i, out):
        o = ops.GemmOut()
        o.out_f32, o.ldo, o.alpha, o.tile_hint = out.data_ptr(), (N // 2 if swiglu else N), 1.0, bn
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
            for i in range(16):
                run(i + c, outs[c])
        graphs.append(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for c in range(concurrent):
        streams[c].wait_event(e0)
        with torch.cuda.stream(streams[c]):
            for _ in range(4):
                graphs[c].replay()
        ev = torch.cuda.Event()
        ev.record(streams[c])
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    n = 64 * concurrent
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:7s} M={M:5d} N={N:5d} K={K:5d} 2cta={two} bn={bn:3d} split={sk} streams={concurrent}: {us:7.2f} us/gemm  "
          f"{2.0 * M * N * K * 3 / us / 1e6:7.1f} TFLOP/s executed", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "layer":
    # the four ViT-L block GEMMs exactly as the engine issues them under the throughput policy
    for conc in (1, 4, 8):
        bench("qkv", 512, 3072, 1024, 0, 1, 1, conc)
        bench("proj", 512, 1024, 1024, 0, 1, 4, conc)
        bench("fc1sw", 512, 5504, 1024, 0, 1, 1, conc, swiglu=True)
        bench("fc2", 512, 1024, 2752, 0, 1, 4, conc)
        bench("proj1", 512, 1024, 1024, 0, 1, 1, conc)
        bench("fc2_1", 512, 1024, 2752, 0, 1, 1, conc)
    sys.exit(0)

if __name__ == "__main__":
    for name in sys.argv[1:] or ["qkv", "fc1", "fc2", "qkv_b4"]:
        M, N, K = SHAPES[name]
        sk = 4 if name in ("proj", "fc2") else 1
        for conc in (1, 4):
            for two, bn in ((0, 128), (0, 256), (1, 128), (1, 256)):
                bench(name, M, N, K, two, bn, sk, conc)

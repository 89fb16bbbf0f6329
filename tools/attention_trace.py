"""Timeline of CTA (0,0,0) of the fused attention kernel (clock64 stamps written through the debug hook
psam_debug_attention_trace; the product path passes a null pointer).  usage: python tools/attention_trace.py [L] [H]"""
import ctypes
import os
import sys
from ctypes import byref

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "point-sam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from psam_b200 import native as nv, ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
TPC = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
dh, B = 64, 1
D = H * dh
qkv = ops.Split(B * L, 3 * D, dev)
qkv.t.normal_()
att = ops.Split(B * L, D, dev)
mk = lambda col: qkv.operand(rows=L, k=dh, col=col, nb1=H, b1_stride=dh, nb2=B, b2_stride=L * qkv.pitch)
qa, ka, va = mk(0), mk(D), mk(2 * D)
run = lambda: nv.check(nv.lib().psam_attention_bf16x3(byref(qa), byref(ka), byref(va), att.ptr(), att.plane, att.pitch, dh,
                                                      L * att.pitch, dh ** -0.5, nv.stream()), "attention")
lib = nv.lib()._real if hasattr(nv.lib(), "_real") else nv.lib()
lib.psam_debug_attention_tiles(TPC)
trace = torch.zeros(256, dtype=torch.int64, device=dev)
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.psam_debug_attention_trace(ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
lib.psam_debug_attention_trace(ctypes.c_void_p(0))
t = trace.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None
print(f"L={L} H={H}: q_full seen {rel(t[1])}, end {rel(t[200])} (clocks since the CTA passed its prologue)")
print("item | S issued | s_full seen | chunk0 done | exps done | p_full arrive | PV issued")
nt = min(TPC, (L + 127) // 128)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    run()
e1.record()
torch.cuda.synchronize()
print(f"tiles per CTA {TPC}: {e0.elapsed_time(e1) / 50 * 1e3:.2f} us per launch (back to back, warm L2)")
G = nt * ((L + 127) // 128)
for g in range(G):
    b = 2 + 8 * g
    print(f"{g:4d} | {rel(t[b])} | {rel(t[b + 2])} | {rel(t[b + 3])} | {rel(t[b + 4])} | {rel(t[b + 5])} | {rel(t[b + 1])}")

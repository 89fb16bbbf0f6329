"""ncu CSV (gpu__time_duration.sum, sm__cycles_active.sum, dram__bytes_*.sum per launch) -> markdown launch table.
usage: python tools/ncu_launch_table.py launches.csv out.md "title" """
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
per = {}
for r in rows:
    e = per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
    e[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)


agg = {}
for p in per.values():
    k = (p["name"], p["grid"], p["block"])
    a = agg.setdefault(k, {"n": 0, "us": 0.0, "cyc": 0.0, "bytes": 0.0})
    a["n"] += 1
    a["us"] += to_us(*p["gpu__time_duration.sum"])
    a["cyc"] += p.get("sm__cycles_active.sum", (0.0, ""))[0]
    a["bytes"] += to_bytes(*p.get("dram__bytes_read.sum", (0, "byte"))) + to_bytes(*p.get("dram__bytes_write.sum", (0, "byte")))
tot_us = sum(a["us"] for a in agg.values())
tot_cyc = sum(a["cyc"] for a in agg.values()) or 1.0
out = [f"# {sys.argv[3]}", "",
       f"{sum(a['n'] for a in agg.values())} launches, {tot_us:.0f} us total (cold-cache, serialised per-launch times: compare SHARES). "
       "`SM-time share` = share of sm__cycles_active.sum (SM x cycles actually occupied) - what a kernel costs when other "
       "clouds run concurrently.", "",
       "| kernel | grid | block | launches | total us | avg us | time share | SM-time share | DRAM MB/launch |", "|---|---|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append(f"| {k[0][:70]} | {k[1]} | {k[2]} | {a['n']} | {a['us']:.1f} | {a['us'] / a['n']:.1f} | {100 * a['us'] / tot_us:.1f}% | "
               f"{100 * a['cyc'] / tot_cyc:.1f}% | {a['bytes'] / a['n'] / 1e6:.2f} |")
fam = {}
for k, a in agg.items():
    n = k[0]
    f_ = ("GEMM (tcgen05)" if "gemm_tc" in n else "fused attention (tcgen05)" if ("attention_tc" in n or "attention_pair" in n or "attention_flow" in n) else "FPS" if "fps" in n else
          "kNN / gathers / interp" if ("knn" in n or "gather" in n or "border" in n) else "LayerNorm family" if ("layernorm" in n or "swiglu" in n or "interp_ln" in n) else
          "decoder SIMT (linear / small attention)" if ("linear" in n or "attention_small" in n or "decoder" in n) else "other")
    b = fam.setdefault(f_, [0.0, 0.0, 0])
    b[0] += a["us"]
    b[1] += a["cyc"]
    b[2] += a["n"]
out += ["", "| family | launches | time share | SM-time share |", "|---|---|---|---|"]
for f_, b in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {f_} | {b[2]} | {100 * b[0] / tot_us:.1f}% | {100 * b[1] / tot_cyc:.1f}% |")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[-len(fam) - 3:]))

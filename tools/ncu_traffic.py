"""Summarise an ncu CSV (dram bytes + duration per launch) for the GEMM kernels -> profiles/rNN_gemm_traffic.json"""
import csv, json, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
per = {}
for r in rows:
    per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
g = [p for p in per.values() if "gemm_tc" in p["name"]]  # one-shot, 2-CTA and persistent tcgen05 GEMM kernels
rd = sum(to_bytes(*p["dram__bytes_read.sum"]) for p in g)
wr = sum(to_bytes(*p["dram__bytes_write.sum"]) for p in g)
us = sum(to_us(*p["gpu__time_duration.sum"]) for p in g)
allus = sum(to_us(*p["gpu__time_duration.sum"]) for p in per.values())
out = {"kernel": "gemm_tc_*_kernel", "how": "ncu launch list of one eager step, throughput tile policy (tools/r2_evidence.sh)", "launches": len(g), "dram_bytes_per_launch": (rd + wr) / len(g), "dram_read_bytes_total": rd,
       "dram_write_bytes_total": wr, "gemm_time_us_total": us, "step_time_us_total": allus, "share_of_step": us / allus,
       "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one eager step of config c2"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)

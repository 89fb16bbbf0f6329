"""ncu report -> short markdown table of the metrics the profiles/ summaries quote.
usage: python tools/ncu_summary.py file.ncu-rep|raw.csv out.md "free text header" [kernel-name-substring]"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]
# a .csv argument is the `--page raw --csv` export made on the GPU box (the .ncu-rep itself can exceed the gpurun_out/ cap)
raw = open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else \
    subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
want = sys.argv[4] if len(sys.argv) > 4 else None
out = []
for r in rows[2:]:
    name = r[head.index("Kernel Name")]
    if want and want not in name:
        continue
    out += [f"# ncu --set full --clock-control none: {name[:110]}", "", sys.argv[3], "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEEP:
        if k in head:
            i = head.index(k)
            out.append(f"| {k} | {r[i]} | {units[i]} |")
    out.append("")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))

"""SASS evidence table: which Blackwell instructions each kernel of libpsam_b200.so contains (cuobjdump -sass).
usage: python tools/sass_summary.py [out.md]      (also called from __graft_entry__.build())
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR,
mbarrier expect-tx/try_wait -> SYNCS, st.async -> ST*.ASYNC / STAS, packed fp32 -> FFMA2/FADD2, redux.sync -> REDUX."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "point-sam_b200", "lib", "libpsam_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"), ("UTMAPF", r"\bUTMAPF|\bUTMACCTL"),
            ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("STAS / st.async", r"\bSTAS|\bST\S*\.ASYNC"),
            ("REDUX", r"\bREDUX|\bCREDUX"), ("FFMA2/FADD2", r"\bFFMA2|\bFADD2"), ("FMNMX3", r"\bFMNMX3"), ("MUFU.EX2", r"\bMUFU\.EX2"),
            ("HMMA (legacy)", r"\bHMMA")]


def main(out_path=None):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                per[cur][name] += 1
    cols = [n for n, _ in PATTERNS]
    tot = collections.Counter()
    rows = []
    for k, c in per.items():
        tot.update(c)
        if any(c[n] for n in cols):
            rows.append("| `" + k[:78] + "` | " + " | ".join(str(c[n]) if c[n] else "" for n in cols) + " |")
    text = ["# SASS evidence (cuobjdump -sass point-sam_b200/lib/libpsam_b200.so, sm_100a)", "",
            "Counts of Blackwell-specific instructions per kernel; regenerated at build time by tools/sass_summary.py.", "",
            "| kernel | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)] + sorted(rows) + [
            "| **total** | " + " | ".join(str(tot[n]) for n in cols) + " |", ""]
    if out_path:
        open(out_path, "w").write("\n".join(text))
    return tot


if __name__ == "__main__":
    t = main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_sass_summary.md"))
    print(dict(t))

"""Top stall sites of an `ncu --page source --csv` export.  usage: python tools/ncu_hot.py source.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
head = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
iS, iA, iI = head.index("# Samples"), head.index("Address"), head.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(head) if h.startswith("stall_")]
body = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in body) or 1
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][iS]))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]) if r[c].isdigit() else 0, head[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[iS]) * 100.0 / tot:5.1f}%  exec={r[iI]:>7s}  {r[1].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")

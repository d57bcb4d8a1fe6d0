"""Summarise an `ncu --page source --csv` export: samples by opcode and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv > x.csv; python profiles/analyze_source.py x.csv [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = []
for r in rows[h + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"):
        break            # the next launch's section: one launch is enough
    if len(r) == len(hdr):
        body.append(r)
col = {n: hdr.index(n) for n in ("# Samples", "Source", "Instructions Executed", "stall_long_sb", "stall_barrier",
                                 "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected")}
tot = sum(int(r[col["# Samples"]]) for r in body)
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
print("samples", tot, "sass lines", len(body), "warp instr executed", sum(int(r[col["Instructions Executed"]]) for r in body))
for k in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected"):
    print("  %-20s %5.1f %%" % (k, 100.0 * sum(int(r[col[k]]) for r in body) / max(tot, 1)))
c, ex = Counter(), Counter()
for r in body:
    parts = r[col["Source"]].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += int(r[col["# Samples"]])
    ex[op] += int(r[col["Instructions Executed"]])
print("by opcode (samples %, executed):")
for op, n in c.most_common(12):
    print("  %-10s %5.1f %%  %d" % (op, 100.0 * n / max(tot, 1), ex[op]))
print("hottest lines:")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][col["# Samples"]]))[:top_n]:
    print("  #%-5d %6s  %-60s long=%s bar=%s short=%s" % (i, r[col["# Samples"]], r[col["Source"]].strip()[:60],
          r[col["stall_long_sb"]], r[col["stall_barrier"]], r[col["stall_short_sb"]]))

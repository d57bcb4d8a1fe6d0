"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python profiles/analyze_cuda_lines.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = []
fname, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if out and any(o[0] == "#fn" for o in out[-1:]):
            pass
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":     # a CUDA source line (SASS rows carry an address)
        try:
            smp = int(r[hdr.index("# Samples")])
            ex = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        g = lambda k: int(r[hdr.index(k)] or 0) if k in hdr else 0
        out.append((smp, ex, fname, r[0], r[1].strip()[:90], g("stall_long_sb"), g("stall_barrier"), g("stall_short_sb"), g("stall_wait")))
tot = sum(o[0] for o in out)
print("total samples on CUDA lines:", tot, " warp instr:", sum(o[1] for o in out))
for o in sorted(out, key=lambda t: -t[0])[:top_n]:
    print("%5.1f%% %9d  %s:%s  long=%d bar=%d short=%d wait=%d | %s" % (100.0 * o[0] / max(tot, 1), o[1], o[2], o[3], o[5], o[6], o[7], o[8], o[4]))

"""Summaries of ncu reports for profiles/: per-kernel key metrics (raw page) and the launch list.
usage: python profiles/summarize.py metrics X.ncu-rep > out.txt ; python profiles/summarize.py launches launches.csv > out.txt"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio")

if sys.argv[1] == "metrics":
    out = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== %s" % r[kn][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-85s %s %s" % (k, r[i], units[i]))
else:
    rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    hdr = rows[0]
    kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[kn].split("(")[0][-70:]
        tot[name][0] += 1
        tot[name][1] += float(r[val].replace(",", ""))
    s = sum(v[1] for v in tot.values())
    print("launches: %d   total gpu time %.1f us (cold cache, serialised: compare shares)" % (len(rows) - 1, s / 1e3))
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%6.2f %%  %5d x  avg %9.2f us   %s" % (100 * t / s, n, t / n / 1e3, name))

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_training.py tests/test_gpu_nodeclf.py -m gpu -q -x > gpurun_out/r5_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r5_tests.log
GNNSEG_LIB=gnn_fpga_b200/libgnnseg_dtrace.so timeout -k 5 120 python scripts/train_profile.py acts64 2 2>&1 | tail -26
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_launches.csv python scripts/train_profile.py acts64 2 > gpurun_out/r5_ncu.log 2>&1
python - <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open("gpurun_out/r5_launches.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); iu=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ik].split("(")[0][:60]; t=float(r[iv].replace(",",""))
    if r[iu]=="ns": t/=1000
    elif r[iu]=="ms": t*=1000
    agg.setdefault(name,[]).append(t)
for k,x in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    if "tc_" in k: print("  %-62s n=%3d  sum %8.1f  last %7.1f"%(k,len(x),sum(x),x[-1]))
PY

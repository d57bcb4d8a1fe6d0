#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 1; do
GNNSEG_DPROP_MODE=$m timeout -k 5 150 python -m pytest tests/test_gpu_training.py tests/test_gpu_nodeclf.py -m gpu -q -x > gpurun_out/r5_tests_m$m.log 2>&1; echo "tests mode $m rc $?"; tail -2 gpurun_out/r5_tests_m$m.log
GNNSEG_DPROP_MODE=$m timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_launches_m$m.csv python scripts/train_profile.py ${WL:-acts64} 2 > gpurun_out/r5_ncu_m$m.log 2>&1
python - $m <<'PY'
import csv,sys,collections
v=sys.argv[1]
rows=[r for r in csv.reader(open("gpurun_out/r5_launches_m%s.csv"%v)) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); iu=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ik].split("(")[0][:60]; t=float(r[iv].replace(",",""))
    if r[iu]=="ns": t/=1000
    elif r[iu]=="ms": t*=1000
    agg.setdefault(name,[]).append(t)
tot=sum(sum(x) for x in agg.values())
print("== mode", v, "total us %.0f over %d launches"%(tot, sum(len(x) for x in agg.values())))
for k,x in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    if "tc_kernel" in k: print("  %-62s n=%3d  sum %8.1f  last %7.1f"%(k,len(x),sum(x),x[-1]))
PY
done

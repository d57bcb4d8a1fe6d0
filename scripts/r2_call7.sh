#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ro in none always; do
  GNNSEG_BENCH_REORDER=$ro timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2d_acts64_$ro.json 2> gpurun_out/r2d_acts64_$ro.err
  GNNSEG_BENCH_REORDER=$ro timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2d_mu200_$ro.json 2> gpurun_out/r2d_mu200_$ro.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("e2e",{}).get("stages_ms"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-mu200"
GNNSEG_BENCH_REORDER=always timeout 300 $CMD > gpurun_out/r2_ncu_plain2.log 2>&1 && \
GNNSEG_BENCH_REORDER=always timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_gather|edge_final" -s 6 -c 3 -o gpurun_out/r2_fused_acts64_reorder -f $CMD > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log

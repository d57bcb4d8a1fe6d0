#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ro in none always; do for win in 4096 256 64 32; do
  GNNSEG_ORDER_WINDOW=$win GNNSEG_BENCH_REORDER=$ro timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 --no-e2e > gpurun_out/r2e_acts64_${ro}_w$win.json 2> gpurun_out/r2e_acts64_${ro}_w$win.err
  GNNSEG_ORDER_WINDOW=$win GNNSEG_BENCH_REORDER=$ro timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2e_mu200_${ro}_w$win.json 2> gpurun_out/r2e_mu200_${ro}_w$win.err
done; done
for ro in none always; do
  GNNSEG_FUSED_CFG=10 GNNSEG_BENCH_REORDER=$ro timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 --no-e2e > gpurun_out/r2e_acts64_${ro}_noorder.json 2> gpurun_out/r2e_acts64_${ro}_noorder.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-300:])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-mu200"
GNNSEG_ORDER_WINDOW=64 GNNSEG_BENCH_REORDER=always timeout 300 $CMD > gpurun_out/r2_ncu_plain3.log 2>&1 && \
GNNSEG_ORDER_WINDOW=64 GNNSEG_BENCH_REORDER=always timeout 900 ncu --set full --clock-control none -k regex:"fused_gather" -s 2 -c 1 -o gpurun_out/r2_fused_acts64_reorder_w64 -f $CMD > gpurun_out/r2_ncu3.log 2>&1
tail -2 gpurun_out/r2_ncu3.log

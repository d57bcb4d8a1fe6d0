#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0 1; do
GNNSEG_PDL=$p timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2p_acts64_pdl$p.json 2> gpurun_out/r2p_acts64_pdl$p.err
GNNSEG_PDL=$p timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2p_mu200_pdl$p.json 2> gpurun_out/r2p_mu200_pdl$p.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), d["e2e"]["stages_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-mu200"
timeout -k 10 300 $CMD > gpurun_out/r2_ncu_plain5.log 2>&1 && \
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_acts64_e2e.csv $CMD > gpurun_out/r2_ncu5.log 2>&1
tail -2 gpurun_out/r2_ncu5.log | cut -c1-300

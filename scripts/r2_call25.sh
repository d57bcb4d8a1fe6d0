#!/bin/bash
# how much does the fused gather lose when the SM's shared-memory carve-out is what the MLP kernel needs?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sm in 0 16384 38400 57344; do
for w in acts64 mu200; do
GNNSEG_GATHER_SMEM=$sm timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2y_${w}_smem$sm.json 2> gpurun_out/r2y_${w}_smem$sm.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2y_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

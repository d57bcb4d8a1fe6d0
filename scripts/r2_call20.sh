#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in "0 0" "1 0" "0 1"; do set -- $m
for w in acts64 mu200; do
GNNSEG_PDL_MLP=$1 GNNSEG_PDL_GATHER=$2 timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2t_${w}_mlp$1_g$2.json 2> gpurun_out/r2t_${w}_mlp$1_g$2.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

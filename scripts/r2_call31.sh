#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1 2; do for m in 2 1; do
GNNSEG_STREAM_ASSEMBLE=$m timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r3d_acts64_asm${m}_$r.json 2> gpurun_out/r3d_acts64_asm${m}_$r.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

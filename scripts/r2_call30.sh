#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 2 1; do for o in 1 0; do
GNNSEG_STREAM_ASSEMBLE=$m GNNSEG_ASSEMBLE_OWN_STREAM=$o timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r3c_mu200_asm${m}_own$o.json 2> gpurun_out/r3c_mu200_asm${m}_own$o.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3c_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests19.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests19.log
for w in acts64 mu200 toy2d; do
timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2u_$w.json 2> gpurun_out/r2u_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2u_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

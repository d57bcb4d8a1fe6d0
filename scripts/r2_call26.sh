#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_store.py -m gpu -x -q > gpurun_out/r2_tests21.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests21.log
for r in 1 2; do for p in 1 0; do
GNNSEG_ASSEMBLE_OWN_STREAM=$p timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2z_acts64_own${p}_$r.json 2> gpurun_out/r2z_acts64_own${p}_$r.err
GNNSEG_ASSEMBLE_OWN_STREAM=$p timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2z_mu200_own${p}_$r.json 2> gpurun_out/r2z_mu200_own${p}_$r.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

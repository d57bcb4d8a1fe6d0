#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests16.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests16.log
for m in 0 1 2; do
GNNSEG_STREAM_ASSEMBLE=$m timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2q_acts64_asm$m.json 2> gpurun_out/r2q_acts64_asm$m.err
GNNSEG_STREAM_ASSEMBLE=$m timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2q_mu200_asm$m.json 2> gpurun_out/r2q_mu200_asm$m.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()}, "tuples", round(d["e2e"]["from_tuples"]["ms_per_step"],2), "blocking", round(d["e2e"]["blocking"]["ms_per_step"],2))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

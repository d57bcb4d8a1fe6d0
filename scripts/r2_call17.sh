#!/bin/bash
# round-2 bench lines for profiles/r2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python bench.py > gpurun_out/r2_line_default.json 2> gpurun_out/r2_line_default.err
timeout -k 10 300 python bench.py --workload toy2d --no-mu200 > gpurun_out/r2_line_toy2d.json 2> gpurun_out/r2_line_toy2d.err
timeout -k 10 300 python bench.py --workload acts64_masked --no-mu200 --no-cpu-baseline > gpurun_out/r2_line_masked.json 2> gpurun_out/r2_line_masked.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_line_reference.json 2> gpurun_out/r2_line_reference.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_line_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g"%d["value"], "ms", d.get("ms_per_step"), "e2e", d.get("e2e",{}).get("ms_per_step"), "train", d.get("train_step",{}).get("ms"), "mu200", d.get("mu200",{}).get("ms_per_step"), d.get("mu200",{}).get("e2e",{}).get("ms_per_step"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

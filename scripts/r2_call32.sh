#!/bin/bash
# cp.async-staged fused gather (GNNSEG_FUSED_CFG 20..34) against the register build: parity, then per-kernel times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 122 331; do
GNNSEG_FUSED_CFG=$c timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r4b_tests_cfg$c.log 2>&1; echo "tests cfg $c rc $?"; tail -2 gpurun_out/r4b_tests_cfg$c.log
done
for c in 0 120 122 131 220 222 231 320 331; do for w in acts64 mu200; do
GNNSEG_FUSED_CFG=$c timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r4b_${w}_cfg$c.json 2> gpurun_out/r4b_${w}_cfg$c.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r4b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r2_tests17.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r2_tests17.log
GNNSEG_FUSED_CFG=20 timeout -k 10 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests18.log 2>&1; echo "tests cfg20 rc $?"; tail -2 gpurun_out/r2_tests18.log
for c in 0 20 21 22 23; do
GNNSEG_FUSED_CFG=$c timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2r_acts64_cfg$c.json 2> gpurun_out/r2r_acts64_cfg$c.err
GNNSEG_FUSED_CFG=$c timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2r_mu200_cfg$c.json 2> gpurun_out/r2r_mu200_cfg$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

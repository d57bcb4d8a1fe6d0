#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_tests22.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests22.log
GNNSEG_EDGE_FINAL=4 timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r2_tests23.log 2>&1; echo "tests ept4 rc $?"; tail -2 gpurun_out/r2_tests23.log
for e in 8 4; do for w in acts64 mu200; do
GNNSEG_EDGE_FINAL=$e timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r3a_${w}_ept$e.json 2> gpurun_out/r3a_${w}_ept$e.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3a_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

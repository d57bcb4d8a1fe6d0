#!/bin/bash
# warp-specialised hidden_dim 64 MLP kernel (gnnseg_mlp_pipe.cu): tests, then A/B against the serial kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "tests rc $?"; tail -15 gpurun_out/r2_tests7.log
for s in pipe serial; do
  GNNSEG_MLP64=$s timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2g_mu200_$s.json 2> gpurun_out/r2g_mu200_$s.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

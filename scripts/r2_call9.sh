#!/bin/bash
# resident weight halves in the hidden_dim 64 MLP kernel: tests, then A/B against the streamed form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests6.log
for s in 0 1; do
  GNNSEG_MLP64_STREAM=$s timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2f_mu200_stream$s.json 2> gpurun_out/r2f_mu200_stream$s.err
done
timeout 600 python bench.py > gpurun_out/r2f_default.json 2> gpurun_out/r2f_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", d.get("e2e",{}).get("ms_per_step"))
        if "mu200" in d: print("  mu200", d["mu200"]["ms_per_step"], d["mu200"]["kernel_ms"], d["mu200"].get("e2e",{}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY

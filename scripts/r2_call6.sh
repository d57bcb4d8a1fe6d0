#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests5.log 2>&1; tail -6 gpurun_out/r2_tests5.log
timeout 300 python scripts/e2e_diag.py acts64 > gpurun_out/r2_e2e_diag.log 2>&1; cat gpurun_out/r2_e2e_diag.log
for cfg in 0 1 2; do
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 --no-e2e > gpurun_out/r2c_acts64_cfg$cfg.json 2> gpurun_out/r2c_acts64_cfg$cfg.err
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2c_mu200_cfg$cfg.json 2> gpurun_out/r2c_mu200_cfg$cfg.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY

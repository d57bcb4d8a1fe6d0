#!/bin/bash
# GPU call 5: fused path v2 (padded adjacency, packed math, balanced order, one-call batch driver)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_store.py -x -q > gpurun_out/r2_fused_tests2.log 2>&1; tail -30 gpurun_out/r2_fused_tests2.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests4.log 2>&1; tail -12 gpurun_out/r2_tests4.log
for cfg in 0 1 2 3 10; do
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 --no-e2e > gpurun_out/r2b_acts64_cfg$cfg.json 2> gpurun_out/r2b_acts64_cfg$cfg.err
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2b_mu200_cfg$cfg.json 2> gpurun_out/r2b_mu200_cfg$cfg.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_default.json 2> gpurun_out/r2b_default.err
tail -c 600 gpurun_out/r2b_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("e2e",{}).get("stages_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY

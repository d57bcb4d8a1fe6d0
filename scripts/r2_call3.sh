#!/bin/bash
# GPU call 3 of round 2: fused inference path — parity first, then timing A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_fused_tests.log 2>&1; tail -30 gpurun_out/r2_fused_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests2.log 2>&1; tail -15 gpurun_out/r2_tests2.log
for cfg in 0 1 2 3; do
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2_bench_fused_cfg$cfg.json 2> gpurun_out/r2_bench_fused_cfg$cfg.err
  GNNSEG_FUSED_CFG=$cfg timeout 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2_bench_mu200_cfg$cfg.json 2> gpurun_out/r2_bench_mu200_cfg$cfg.err
done
GNNSEG_EXACT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2_bench_exact.json 2> gpurun_out/r2_bench_exact.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 400 gpurun_out/r2_bench_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "e2e", d.get("e2e",{}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests15.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/r2_tests15.log
GNNSEG_LIB=gnn_fpga_b200/libgnnseg_trace.so timeout -k 10 200 python scripts/pipe_trace.py 0 32 > gpurun_out/pipe_trace32_v1.txt 2>&1; tail -42 gpurun_out/pipe_trace32_v1.txt
for p in 1 0; do
GNNSEG_MLP_PIPE=$p timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2o_acts64_pipe$p.json 2> gpurun_out/r2o_acts64_pipe$p.err
done
timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2o_mu200.json 2> gpurun_out/r2o_mu200.err
timeout -k 10 300 python bench.py --workload toy2d --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e > gpurun_out/r2o_toy2d.json 2> gpurun_out/r2o_toy2d.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in b200 storeLAST storeNORMAL; do
for w in acts64 mu200; do
GNNSEG_LIB=gnn_fpga_b200/libgnnseg_$v.so timeout -k 10 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2x_${w}_$v.json 2> gpurun_out/r2x_${w}_$v.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2x_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

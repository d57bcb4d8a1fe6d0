#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_store.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2_tests24.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_tests24.log
for r in 1 2; do
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r3b_acts64_$r.json 2> gpurun_out/r3b_acts64_$r.err
timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r3b_mu200_$r.json 2> gpurun_out/r3b_mu200_$r.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-mu200"
timeout -k 10 300 $CMD > gpurun_out/r2_ncu_plain6.log 2>&1 && \
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"assemble|adjacency|order" -c 60 --csv --log-file gpurun_out/r2_launches_assembly.csv $CMD > gpurun_out/r2_ncu6.log 2>&1
python profiles/summarize.py launches gpurun_out/r2_launches_assembly.csv | head

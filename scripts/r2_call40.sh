#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in tc; do
GNNSEG_DENSE_BWD=$v timeout -k 10 300 python bench.py --workload acts64 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-mu200 > gpurun_out/r6_train_$v.json 2> gpurun_out/r6_train_$v.err
done
timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-mu200 > gpurun_out/r6_train_mu200.json 2> gpurun_out/r6_train_mu200.err
GNNSEG_DENSE_BWD=simt timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-mu200 > gpurun_out/r6_train_mu200_simt.json 2> gpurun_out/r6_train_mu200_simt.err
python - <<'PY'
import json
for f in ["r6_train_tc","r6_train_mu200","r6_train_mu200_simt"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        ts=d.get("train_step") or {}
        print(f, "fwd ms %.4f"%d["ms_per_step"], "train ms", ts.get("ms"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-600:])
PY

#!/bin/bash
# full GPU suite + default bench line + training step A/B after the tcgen05 dense backward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r6_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r6_tests.log
timeout -k 10 600 python bench.py > gpurun_out/r6_bench_default.json 2> gpurun_out/r6_bench_default.err; echo "bench rc $?"
for v in simt w; do
GNNSEG_DENSE_BWD=$v timeout -k 10 300 python bench.py --workload acts64 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-mu200 > gpurun_out/r6_train_$v.json 2> gpurun_out/r6_train_$v.err
done
python - <<'PY'
import json
for f in ["r6_bench_default","r6_train_simt","r6_train_w"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        ts=d.get("train_step") or {}
        print(f, "fwd ms %.4f"%d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "train ms", ts.get("ms"), "mu200", {k:(d.get("mu200") or {}).get(k) for k in ("ms_per_step",)}, (d.get("mu200") or {}).get("train_step"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-600:])
PY

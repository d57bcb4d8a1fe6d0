#!/bin/bash
# final state of round 2: full GPU suite, smoke(), the default bench line, the reference arm, the masked twin
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q > gpurun_out/r7_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r7_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r7_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r7_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r7_bench_default.json 2> gpurun_out/r7_bench_default.err; echo "bench rc $?"
timeout -k 10 600 python bench.py --workload acts64_masked --steps 50 --no-cpu-baseline > gpurun_out/r7_bench_masked.json 2> gpurun_out/r7_bench_masked.err; echo "masked rc $?"
python - <<'PY'
import json
for f in ["r7_bench_default","r7_bench_masked"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        ts=d.get("train_step") or {}
        mu=d.get("mu200") or {}
        print(f, "fwd ms %.4f"%d["ms_per_step"], "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "e2e val", (d.get("e2e") or {}).get("value"), "train ms", ts.get("ms"), "frac", (d.get("roofline") or {}).get("frac"), "mu200 ms", mu.get("ms_per_step"), "mu200 e2e", (mu.get("e2e") or {}).get("ms_per_step"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-600:])
PY

#!/bin/bash
# what the driver runs at round end, on one box: GPU tests, smoke(), the default bench line, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_final_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2_final_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc $?"
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err; echo "reference rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("acts64 value %.4g ms %.4f roofline %.3f fwd %.3f traffic %s e2e %.4f ms (%.4g) train %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_forward"]["frac"], d["roofline"]["traffic"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["train_step"]["ms"]))
m=d["mu200"]; print("mu200 value %.4g ms %.4f roofline %.3f fwd %.3f e2e %.4f" % (m["value"], m["ms_per_step"], m["roofline"]["frac"], m["roofline_forward"]["frac"], m["e2e"]["ms_per_step"]))
print(d["kernel_ms"], m["kernel_ms"], d["clocks"], d["cpu_baseline"]["value"], d["e2e"]["from_tuples"]["ms_per_step"], d["e2e"]["blocking"]["ms_per_step"])
r=json.loads(open("gpurun_out/r2_final_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["kind"])
PY

#!/bin/bash
# does an L2-sized event group run faster per event? (acts64 kernels on 8 / 16 / 32 / 64 / 128 events)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 8 16 32 64 128; do
GNNSEG_BENCH_BATCH=$b timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-e2e --no-mu200 > gpurun_out/r2s_acts_b$b.json 2> gpurun_out/r2s_acts_b$b.err
done
python - <<'PY'
import json,glob
for b in (8,16,32,64,128):
    f="gpurun_out/r2s_acts_b%d.json"%b
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(b, "ms %.4f"%d["ms_per_step"], "us/event %.2f"%(d["ms_per_step"]*1e3/b), {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, {k:round(v*1e3/b,2) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

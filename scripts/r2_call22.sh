#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1 2; do for p in d 0; do
if [ $p = d ]; then unset GNNSEG_PDL; else export GNNSEG_PDL=0; fi
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-mu200 > gpurun_out/r2v_acts64_pdl${p}_$r.json 2> gpurun_out/r2v_acts64_pdl${p}_$r.err
timeout -k 10 300 python bench.py --workload mu200 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r2v_mu200_pdl${p}_$r.json 2> gpurun_out/r2v_mu200_pdl${p}_$r.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms %.4f"%d["ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["e2e"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY

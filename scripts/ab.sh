#!/bin/bash
# A/B of scheduling switches: prints ms_per_step and per-kernel ms for each environment setting given as arguments
# usage: scripts/ab.sh "VAR=a VAR2=b" "VAR=c" ...
for envs in "$@"; do
  out=$(env $envs python bench.py --steps 40 --no-e2e --no-train --no-cpu-baseline ${WL:+--workload $WL} 2>/dev/null | tail -1)
  python - "$envs" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print("%-50s step %.4f ms  kernels %s" % (sys.argv[1], d["ms_per_step"], {k: round(v, 4) for k, v in d["kernel_ms"].items()}))
PY
done

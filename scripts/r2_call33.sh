#!/bin/bash
# tcgen05 weight-gradient kernel (gnnseg_wgrad_tc.cu): training parity, then the training step with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_training.py tests/test_gpu_nodeclf.py -m gpu -q -x > gpurun_out/r5_tests.log 2>&1; echo "tests rc $?"; tail -15 gpurun_out/r5_tests.log
for v in tc simt; do
GNNSEG_DENSE_BWD=$v timeout -k 10 300 python bench.py --workload acts64 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-mu200 > gpurun_out/r5_train_$v.json 2> gpurun_out/r5_train_$v.err
python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r5_train_%s.json"%v).read().strip().splitlines()[-1])
    print(v, "fwd ms %.4f"%d["ms_per_step"], "train_step", d.get("train_step"))
except Exception as e:
    print(v, "ERR", e, open("gpurun_out/r5_train_%s.err"%v).read()[-800:])
PY
done

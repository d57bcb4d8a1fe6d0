#!/bin/bash
# 8 ranks: e2e with and without binding every rank to the cores next to its GPU (pinned memory on the GPU's NUMA node)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
for b in 1 0; do
GNNSEG_BIND_CPUS=$b timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 8 --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_8gpu_bind$b.json 2> gpurun_out/r2_bench_8gpu_bind$b.err
done
python - <<'PY'
import json
for b in (1,0):
    try:
        d=json.loads(open("gpurun_out/r2_bench_8gpu_bind%d.json"%b).read().strip().splitlines()[-1])
        print("bind",b, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["stages_ms"], d.get("rank0_cpu_binding"))
    except Exception as e:
        print("bind",b,"ERR",e, open("gpurun_out/r2_bench_8gpu_bind%d.err"%b).read()[-500:])
PY
head -14 gpurun_out/r2_topo.txt

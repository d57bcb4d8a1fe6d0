#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r7_bench_8gpu.json 2> gpurun_out/r7_bench_8gpu.err; echo "rc $?"
tail -1 gpurun_out/r7_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scores_allgather_ms','value_with_gather','gathered_scores_bit_equal_to_local_recompute')}, (d.get('e2e') or {}).get('ms_per_step'), (d.get('train_step') or {}).get('ms'))"

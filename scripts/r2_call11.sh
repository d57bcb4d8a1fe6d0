#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --workload mu200 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train"
timeout -k 10 300 $CMD > gpurun_out/r2_ncu_plain4.log 2>&1 && \
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"pipe64" -s 4 -c 2 -o gpurun_out/r2_pipe64_mu200 -f $CMD > gpurun_out/r2_ncu4.log 2>&1
tail -3 gpurun_out/r2_ncu4.log

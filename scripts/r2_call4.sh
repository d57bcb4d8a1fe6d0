#!/bin/bash
# GPU call 4: re-test fixes, then ncu of the fused kernels (acts64)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests3.log 2>&1; tail -8 gpurun_out/r2_tests3.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-mu200"
timeout 300 $CMD > gpurun_out/r2_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_gather|edge_final|node_mlp_kernel_tc" -s 12 -c 6 -o gpurun_out/r2_fused_acts64 -f $CMD > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log

#!/bin/bash
# round-2 evidence: launch lists and full captures of the two dominant kernels, both workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in acts64 mu200; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-mu200"
  timeout -k 10 300 $CMD > gpurun_out/r2_final_plain_$w.log 2>&1 || { echo "plain run failed $w"; continue; }
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_$w.csv $CMD > gpurun_out/r2_final_ncu_l_$w.log 2>&1
  timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"fused_gather|node_mlp_kernel_pipe|edge_final" -s 9 -c 4 -o gpurun_out/r2_final_$w -f $CMD > gpurun_out/r2_final_ncu_f_$w.log 2>&1
  tail -1 gpurun_out/r2_final_ncu_f_$w.log
done

#!/bin/bash
# ncu --set full of the two tcgen05 dense-backward kernels (one launch each, inside a warm training step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|dprop_tc_kernel" -s 8 -c 2 -f -o gpurun_out/r5_dense_tc python scripts/train_profile.py acts64 2 > gpurun_out/r5_ncu_full.log 2>&1
echo rc $?; tail -3 gpurun_out/r5_ncu_full.log; ls -la gpurun_out/r5_dense_tc.ncu-rep

#!/bin/bash
# GPU call 2 of round 2: new store/loader tests, tanh accuracy A/B, contiguous-range gather A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; tail -15 gpurun_out/r2_tests1.log
rm -f gpurun_out/parity_report.jsonl
for lib in b200 tanhfast edgeacc; do
  echo "{\"lib\": \"$lib\"}" >> gpurun_out/parity_report.jsonl
  GNNSEG_LIB=gnn_fpga_b200/libgnnseg_$lib.so python -m pytest tests/test_gpu_fullsize.py -q -k large_weights > gpurun_out/r2_acc_$lib.log 2>&1
done
cat gpurun_out/parity_report.jsonl
for r in 0 1 2; do
  for wl in acts64 mu200; do
    GNNSEG_GATHER_RANGES=$r python scripts/locality_experiment.py $wl none phi layer_phi > gpurun_out/r2_loc2_${wl}_r$r.log 2>&1
  done
done
grep -h '"kernel_ms"' gpurun_out/r2_loc2_*.log | wc -l

#!/bin/bash
# Run the GPU test files one process each (a trapped kernel poisons only its own file), then the probe.
mkdir -p gpurun_out
rm -f gpurun_out/test_diagnostics.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in test_gpu_elementwise test_gpu_gemm test_gpu_bsi_api test_gpu_dit; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log
done
if [ "$1" != "noprobe" ]; then
  timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
  echo "probe exit $?" >> gpurun_out/summary.txt
  cat gpurun_out/probe.log | tail -12
fi

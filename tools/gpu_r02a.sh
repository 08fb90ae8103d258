#!/bin/bash
# Round-2 GPU pass A: full GPU suite (one log per file), kernel timeline with and without PDL, a short bench with the side measurements.
mkdir -p gpurun_out/r02
rm -f gpurun_out/test_diagnostics.txt gpurun_out/full_depth_parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r02/nvsmi.txt 2>&1
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02/$b.log 2>&1
  echo "$b exit $?" | tee -a gpurun_out/r02/summary.txt
  tail -2 gpurun_out/r02/$b.log
done
BSI_PDL=0 timeout 600 python tools/gpu_timeline.py --k 4 > gpurun_out/r02/timeline_pdl0.txt 2>&1; echo "timeline0 exit $?"
BSI_PDL=1 timeout 600 python tools/gpu_timeline.py --k 4 > gpurun_out/r02/timeline_pdl1.txt 2>&1; echo "timeline1 exit $?"
head -3 gpurun_out/r02/timeline_pdl0.txt; head -3 gpurun_out/r02/timeline_pdl1.txt
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02/bench_a.json; tail -5 gpurun_out/r02/bench_a.err
cp gpurun_out/full_depth_parity.jsonl gpurun_out/r02/ 2>/dev/null

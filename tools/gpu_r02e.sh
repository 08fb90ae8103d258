#!/bin/bash
# full GPU suite (one log per file) + smoke + one headline bench line with all side measurements
mkdir -p gpurun_out/r02
rm -f gpurun_out/test_diagnostics.txt gpurun_out/full_depth_parity.jsonl gpurun_out/r02/summary.txt
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02/$b.log 2>&1
  echo "$b exit $? $(tail -1 gpurun_out/r02/$b.log)" | tee -a gpurun_out/r02/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02/smoke.log
timeout 1200 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/r02/bench_e.json 2> gpurun_out/r02/bench_e.err; echo "bench exit $?"
tail -c 4500 gpurun_out/r02/bench_e.json; tail -5 gpurun_out/r02/bench_e.err
cp gpurun_out/full_depth_parity.jsonl gpurun_out/r02/ 2>/dev/null

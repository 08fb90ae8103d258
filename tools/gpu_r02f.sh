#!/bin/bash
# Round-2 final pass on one GPU: full GPU suite, smoke, headline bench + the two other sampling configurations + the reference arm
mkdir -p gpurun_out/r02s
rm -f gpurun_out/test_diagnostics.txt gpurun_out/full_depth_parity.jsonl gpurun_out/r02s/summary.txt
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02s/$b.log 2>&1
  echo "$b exit $? $(tail -1 gpurun_out/r02s/$b.log)" | tee -a gpurun_out/r02s/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02s/smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02s/bench_final.json 2> gpurun_out/r02s/bench_final.err; echo "bench exit $?"
timeout 900 python bench.py --config imagenet32-dit --steps 2 --warmup 3 --no-side > gpurun_out/r02s/bench_in32.json 2>> gpurun_out/r02s/bench_final.err; echo "bench in32 exit $?"
timeout 900 python bench.py --config cifar10-vdm --steps 3 --warmup 3 --no-side > gpurun_out/r02s/bench_cifar.json 2>> gpurun_out/r02s/bench_final.err; echo "bench cifar exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s/bench_ref.json 2>> gpurun_out/r02s/bench_final.err; echo "bench ref exit $?"
for f in bench_final bench_in32 bench_cifar; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02s/$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms", round(d["ms_per_step"]), "tf", round(d["config"]["whole_step_tflops_per_gpu"]), "gemm", round(d["roofline"]["achieved"]), d["clocks"]["sm_mhz"], {k: (round(v["value"],1) if k=="elbo" else round(v["ms_per_step"],1)) for k,v in d.items() if k in ("elbo","train_step")})
PY
done
tail -3 gpurun_out/r02s/bench_final.err
cp gpurun_out/full_depth_parity.jsonl gpurun_out/r02s/ 2>/dev/null

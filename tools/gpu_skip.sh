#!/bin/bash
# Which share of a sampler step is a kernel family worth under the 1 kW power cap?  (timing only: skipped kernels give wrong results)
mkdir -p gpurun_out/r02
for s in none ln attn ln,attn; do
  echo "== BSI_DEBUG_SKIP=$s"
  BSI_DEBUG_SKIP=$s timeout 600 python bench.py --steps 2 --warmup 1 --k 64 --no-side --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', round(d['ms_per_step'],1), 'samples/s', round(d['value'],2), 'clocks', d['clocks'], 'gemm TF', round(d['roofline']['achieved'],1), 'gemm share', round(d['roofline']['gemm_share_of_step'],3))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done

#!/bin/bash
# round 2, pass l: bulk-copy pipelined LayerNorm (inference / ELBO) and gate backward
mkdir -p gpurun_out/r02l
O=gpurun_out/r02l
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_dit.py tests/test_gpu_full_depth.py tests/test_gpu_bsi_api.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee $O/tests.log
(BSI_LN_PIPE=0 timeout 200 python tools/gpu_train_kernels.py | grep "inference"; timeout 200 python tools/gpu_train_kernels.py) 2>&1 | tee $O/train_kernels.jsonl
for p in 0 1; do BSI_LN_PIPE=$p timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_lnpipe$p.json 2> $O/bench_lnpipe$p.err; done
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl
python - <<'PY'
import json
for p in (0, 1):
    try:
        d = json.loads(open(f"gpurun_out/r02l/bench_lnpipe{p}.json").read().strip().splitlines()[-1])
        print("LN_PIPE", p, "samples/s", d["value"], "e2e", d["e2e"]["value"], "elbo", d.get("elbo", {}).get("value"), "train", d.get("train_step", {}).get("ms_per_step"), "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print("bench", p, "failed", e)
PY

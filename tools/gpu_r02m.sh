#!/bin/bash
# round 2, pass m: equal-work (heads + tails) weight-gradient GEMM with a double-buffered accumulator
mkdir -p gpurun_out/r02m
O=gpurun_out/r02m
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee $O/tests.log
PROBE_M=32768 timeout 300 python tools/gpu_wgrad.py 2>&1 | tee $O/wgrad.jsonl
timeout 300 python tools/gpu_wgrad.py 2>&1 | tee -a $O/wgrad.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl

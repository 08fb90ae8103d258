#!/bin/bash
# weight-copy cache across micro-batches: tests + the 1-GPU training step at global batch 1024 (4 micro-batches of 256)
mkdir -p gpurun_out/r02r
O=gpurun_out/r02r
timeout 900 python -m pytest tests/test_gpu_dit_train.py tests/test_gpu_optim.py tests/test_gpu_reference_speed.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4 | tee $O/tests.log
timeout 400 python tools/gpu_train.py --global-batch 1024 --steps 4 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl

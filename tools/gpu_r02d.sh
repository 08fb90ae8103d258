#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_optim.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/gpu_train.py --global-batch 128 --dropout 0.05 2>&1 | tail -1 | tee -a gpurun_out/r02/train_attbwd_ab.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 2>&1 | tail -1 | tee -a gpurun_out/r02/train_attbwd_ab.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --dropout 0.05 --profile 2>&1 | tail -40 > gpurun_out/r02/train_step_kernels.txt

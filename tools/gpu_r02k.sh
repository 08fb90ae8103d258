#!/bin/bash
# round 2, pass k: bulk-copy pipelined LayerNorm kernels of the training path
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_optim.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee $O/tests.log
(BSI_TRAIN_PIPE=0 timeout 200 python tools/gpu_train_kernels.py; timeout 200 python tools/gpu_train_kernels.py) 2>&1 | tee $O/train_kernels.jsonl
for p in 0 1 0 1; do BSI_TRAIN_PIPE=$p timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | sed "s/^{/{\"pipe\": $p, /" | tee -a $O/train.jsonl; done
timeout 300 python tools/gpu_train_timeline.py > $O/timeline.txt 2>&1; grep "train step" $O/timeline.txt

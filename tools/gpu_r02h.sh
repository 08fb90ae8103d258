#!/bin/bash
# round 2, pass h: pipelined attention backward, LayerNorm backward at 2 CTAs/SM, fewer small ops in the DiT backward
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_optim.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -15 | tee $O/tests.log
for v in 1 2 9; do BSI_ATT_BWD_VARIANT=$v timeout 200 python tools/gpu_attbwd.py 2>&1 | tail -2 | tee -a $O/attbwd.jsonl; done
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl
BSI_ATT_BWD_VARIANT=1 timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train_timeline.py 2>&1 | tail -60 | tee $O/train_timeline.txt

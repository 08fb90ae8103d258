#!/bin/bash
# round 2, pass p: 16-bit dropout decisions (one hash per two elements), wider reduce kernel
mkdir -p gpurun_out/r02p
O=gpurun_out/r02p
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_optim.py tests/test_gpu_dit.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -8 | tee $O/tests.log
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train_timeline.py > $O/timeline.txt 2>&1; grep "train step\|attention\|reduce_rows" $O/timeline.txt

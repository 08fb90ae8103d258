#!/bin/bash
# round 2, pass i: MMA issue path of the attention kernels (descriptors hoisted), pipelined backward
mkdir -p gpurun_out/r02i
O=gpurun_out/r02i
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dit_train.py tests/test_gpu_dit.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee $O/tests.log
for v in 1 2 9; do BSI_ATT_BWD_VARIANT=$v timeout 200 python tools/gpu_attbwd.py 2>&1 | tail -1 | tee -a $O/attbwd.jsonl; done
for v in 1 9; do BSI_ATT_VARIANT=$v timeout 300 python tools/gpu_att2.py 2>&1 | tail -1 | tee -a $O/att2.jsonl; done
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 2>&1 | tail -1 | tee -a $O/train.jsonl
timeout 300 python tools/gpu_train.py --global-batch 128 --steps 6 --dropout 0.05 2>&1 | tail -1 | tee -a $O/train.jsonl

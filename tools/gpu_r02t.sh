#!/bin/bash
# 8 GPUs: gradients straight into the arena with the all-reduces started per block during the backward, or as one collective after it
mkdir -p gpurun_out/r02t
O=gpurun_out/r02t
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/gpu_train.py --global-batch 1024 --steps 8 --dropout 0.05 "$@" 2>&1 | grep '^{' | tail -1 | tee -a $O/train_8gpu_overlap.jsonl; }
run --no-overlap
run
run --no-overlap

#!/bin/bash
# A/B inside one box: previous kernels (libbsi_b200_prev.so: LayerNorm backward with register partial sums, 128-query-step attention backward)
# vs the current library, same Python driver
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
BSI_B200_LIB=$PWD/bsi_b200/libbsi_b200_prev.so timeout 300 python tools/gpu_train_timeline.py > $O/timeline_prev.txt 2>&1
timeout 300 python tools/gpu_train_timeline.py > $O/timeline_new.txt 2>&1
BSI_B200_LIB=$PWD/bsi_b200/libbsi_b200_prev.so timeout 300 python tools/gpu_train_timeline.py > $O/timeline_prev2.txt 2>&1
timeout 300 python tools/gpu_train_timeline.py > $O/timeline_new2.txt 2>&1
grep -h "train step\|layernorm_mod_backward\|attention_bwd\|attention_tc2" $O/timeline_*.txt

#!/bin/bash
# ncu --set full of the round-2 training kernels (row-pipelined LayerNorm / gate kernels, weight-gradient GEMM, attention backward)
mkdir -p gpurun_out/r02u
O=gpurun_out/r02u
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"_pipe" -s 12 -c 8 -o $O/prof_train_rows -f python tools/gpu_train_kernels.py > $O/ncu_rows.log 2>&1
PROBE_M=32768 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_wgrad_bf16" -s 3 -c 2 -o $O/prof_wgrad -f python tools/gpu_wgrad.py > $O/ncu_wgrad.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_attention_bwd_tc2|k_attention_dsum" -s 2 -c 2 -o $O/prof_attbwd -f python tools/gpu_attbwd.py > $O/ncu_attbwd.log 2>&1
ls -la $O/*.ncu-rep
CASES="k_layernorm_mod_backward,k_gate_residual_backward_pipe,inference" ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"_pipe" -c 10 -o $O/prof_train_rows2 -f python tools/gpu_train_kernels.py > $O/ncu_rows2.log 2>&1

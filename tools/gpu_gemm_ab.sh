#!/bin/bash
mkdir -p gpurun_out/r02
for v in "" plain4 heavy5 ""; do
  if [ -z "$v" ]; then unset BSI_B200_LIB; else export BSI_B200_LIB=$PWD/bsi_b200/libbsi_b200_$v.so; fi
  timeout 300 python tools/gpu_gemm_ab.py 2>&1 | tail -1 | tee -a gpurun_out/r02/gemm_ab.jsonl
done
unset BSI_B200_LIB
BSI_ATT_VARIANT=9 python tools/gpu_att2.py 2>&1 | tail -3 | tee -a gpurun_out/r02/att2.jsonl

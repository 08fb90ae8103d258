#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dit_train.py tests/test_gpu_backward.py tests/test_gpu_optim.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -6
for v in 0 1 0 1; do
  echo "BSI_TRAIN_FUSED_GELU=$v"
  BSI_TRAIN_FUSED_GELU=$v timeout 300 python tools/gpu_train.py --global-batch 128 --dropout 0.05 2>&1 | tail -1 | tee -a gpurun_out/r02/train_gelu_ab.jsonl
done
HBM_ONE_LAUNCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step_fused|k_q_sample|k_recon_reduce|k_sqerr_reduce|k_patch_operand|k_layernorm_mod" \
  --launch-count 12 -f -o gpurun_out/prof_hbm_r02 python tools/gpu_hbm.py > gpurun_out/r02/ncu_hbm.log 2>&1; echo "ncu hbm exit $?"
BSI_TRAIN_FUSED_GELU=1 timeout 300 python tools/gpu_train.py --global-batch 128 --dropout 0.05 --profile 2>&1 | tail -40 > gpurun_out/r02/train_step_kernels.txt

#!/bin/bash
# ncu evidence of round 2: (1) --set full of the HBM-bound kernels at config-4 sizes, (2) launch list + --set full of one sampler step
mkdir -p gpurun_out/r02
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step_fused|k_q_sample|k_recon_reduce|k_sqerr_reduce|k_patch_operand|k_layernorm_mod" \
  --launch-skip 12 --launch-count 12 -f -o gpurun_out/prof_hbm_r02 python tools/gpu_hbm.py > gpurun_out/r02/ncu_hbm.log 2>&1; echo "ncu hbm exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 1 --warmup 0 --k 2 --no-cpu-baseline --no-side > gpurun_out/r02/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_bf16|k_attention_tc2|k_layernorm_mod" --launch-skip 60 --launch-count 10 -f \
  -o gpurun_out/prof_r02 python bench.py --steps 1 --warmup 0 --k 2 --no-cpu-baseline --no-side > gpurun_out/r02/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3

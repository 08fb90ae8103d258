#!/bin/bash
# Round-end evidence on one B200: all GPU tests, smoke, headline bench, ELBO probe, ncu launch list + full kernel captures.
TAG=${1:-r01}
mkdir -p gpurun_out
bash tools/gpu_round.sh noprobe
bash tools/gpu_one.sh test_gpu_unet test_gpu_optim test_gpu_multi test_gpu_backward test_gpu_dit_train test_gpu_reference_speed
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -c 1800 gpurun_out/bench_$TAG.json
python tools/gpu_elbo.py > gpurun_out/elbo_$TAG.log 2>&1; tail -2 gpurun_out/elbo_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --k 2 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_bf16|k_attention_tc|k_layernorm|k_step_fused|k_patch_operand' -s 8 -c 10 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 0 --k 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
# training-path kernels (config 5): full captures of the weight-gradient GEMM, attention backward, LayerNorm backward and optimizer step
ncu --set full --clock-control none --import-source on -k regex:'k_wgrad|k_attention_bwd|k_layernorm_mod_backward|k_adamw_ema' -s 4 -c 8 -o gpurun_out/prof_train_$TAG -f \
    python tools/gpu_train.py --global-batch 64 --depth 2 --steps 1 > gpurun_out/ncu_train_$TAG.log 2>&1; echo "ncu train exit $?"
python tools/gpu_train.py --global-batch 128 > gpurun_out/train_$TAG.log 2>&1; tail -1 gpurun_out/train_$TAG.log
python tools/gpu_train.py --global-batch 128 --dropout 0.05 >> gpurun_out/train_$TAG.log 2>&1; tail -1 gpurun_out/train_$TAG.log

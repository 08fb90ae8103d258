#!/bin/bash
# 2-GPU pass: the NCCL tests the 1-GPU driver box skips, then bench.py --gpus 2 (multi_gpu_check + sample + elbo + train_step)
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02/test_gpu_multi_2gpu.log 2>&1; echo "multi exit $?"; tail -3 gpurun_out/r02/test_gpu_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02/bench_2gpu.json 2> gpurun_out/r02/bench_2gpu.err; echo "bench2 exit $?"
tail -c 4000 gpurun_out/r02/bench_2gpu.json; tail -8 gpurun_out/r02/bench_2gpu.err

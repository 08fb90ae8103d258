#!/bin/bash
# 8-GPU validation of what the driver will run: bench.py --gpus 8 (multi_gpu_check + sample + sharded elbo + data-parallel train step)
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_bsi_api.py tests/test_gpu_elementwise.py -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02/bench_8gpu.json 2> gpurun_out/r02/bench_8gpu.err; echo "bench8 exit $?"
tail -c 3500 gpurun_out/r02/bench_8gpu.json; tail -5 gpurun_out/r02/bench_8gpu.err

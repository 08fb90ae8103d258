#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each of the DiT GEMM epilogue kinds + attention + layernorm)
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_bf16|k_attention_tc|k_layernorm' -s 8 -c 8 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 0 --k 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -5

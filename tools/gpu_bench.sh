#!/bin/bash
# Round benchmark + ncu evidence on one B200.  Usage: tools/gpu_bench.sh <tag> [bench args]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
# launch list (cold-cache, serialised: compare SHARES): skip the ~150 weight-packing casts, take two denoiser forwards
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --k 2 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches exit $?"
# full capture of the dominant kernel (3 launches of the QKV / MLP GEMMs)
ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16 -s 6 -c 4 -o gpurun_out/prof_gemm_$TAG -f \
    python bench.py --steps 1 --warmup 0 --k 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -12

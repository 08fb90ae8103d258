#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "attention" 2>&1 | tail -15
for v in 0 1 9; do BSI_ATT_BWD_VARIANT=$v timeout 200 python tools/gpu_attbwd.py 2>&1 | tail -2 | tee -a gpurun_out/r02/attbwd.jsonl; done

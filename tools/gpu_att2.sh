#!/bin/bash
mkdir -p gpurun_out/r02
for v in 0 1 2 9; do
  BSI_ATT_VARIANT=$v timeout 300 python tools/gpu_att2.py 2>&1 | tail -2 | tee -a gpurun_out/r02/att2.jsonl
done
echo "== tests variant 1"
BSI_ATT_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_backward.py tests/test_gpu_dit.py tests/test_gpu_dit_train.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5

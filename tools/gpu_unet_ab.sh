#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_full_depth.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "unet or conv or groupnorm or attention_d128" 2>&1 | tail -8
for v in 0 1 0 1; do
  echo "BSI_GN_EPILOGUE=$v"
  BSI_GN_EPILOGUE=$v timeout 300 python tools/gpu_probe_unet.py 2>&1 | tail -1 | tee -a gpurun_out/r02/unet_gn_ab.jsonl
done

#!/bin/bash
# run selected GPU test files, one process each
mkdir -p gpurun_out
for f in "$@"; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/$f.log
done

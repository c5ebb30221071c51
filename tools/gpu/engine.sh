#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_unet_gpu.py tests/test_search_gpu.py; do
  name=$(basename $t .py)
  timeout -k 10 900 python -m pytest $t -q -m gpu -p no:cacheprovider -s > "gpurun_out/${name}.log" 2>&1
  echo "[$t] exit $?"
  tail -n 60 "gpurun_out/${name}.log"
done

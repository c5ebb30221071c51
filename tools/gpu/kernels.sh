#!/bin/bash
# Run each kernel-test group in its own process (a hung kernel must not block the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "sampler_kernels_bit_exact or scorer_argmax or candidates_and_gather or linear" "groupnorm" "conv_gemm" "fused_skip or small_n or im2col or qkv" "attention"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-30)
  timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider > "gpurun_out/k_${name}.log" 2>&1
  echo "group [$grp] exit $?"
  tail -n 25 "gpurun_out/k_${name}.log"
done

#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "operand_path" > gpurun_out/c31_xf_tests.log 2>&1; echo "xf tests rc=$?"; tail -3 gpurun_out/c31_xf_tests.log
timeout -k 5 600 python tools/profile_ops.py 64 --csv gpurun_out/c31_ops.csv > gpurun_out/c31_profile_ops.log 2>&1
sed -n 1,8p gpurun_out/c31_profile_ops.log
grep -E "enc.32x32_block0.norm2\+qkv|enc.16x16_block0.norm2\+qkv|enc.8x8_block0.norm2\+qkv" gpurun_out/c31_ops.csv

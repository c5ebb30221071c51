#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "operand_path" > gpurun_out/c28_xf_tests.log 2>&1; echo "xf tests rc=$?"; tail -15 gpurun_out/c28_xf_tests.log
timeout -k 5 900 python -m pytest tests/test_unet_gpu.py tests/test_search_gpu.py tests/test_classifier_gpu.py -x -q -p no:cacheprovider > gpurun_out/c28_unet_tests.log 2>&1; echo "unet tests rc=$?"; tail -4 gpurun_out/c28_unet_tests.log
timeout -k 5 600 python tools/profile_ops.py 64 --csv gpurun_out/c28_ops_b64.csv > gpurun_out/c28_profile_ops.log 2>&1; echo "profile rc=$?"; head -9 gpurun_out/c28_profile_ops.log
B200NS_FUSED_NORM_A=0 timeout -k 5 600 python tools/profile_ops.py 64 > gpurun_out/c28_profile_ops_unfused.log 2>&1; head -9 gpurun_out/c28_profile_ops_unfused.log
grep -E "qkv" gpurun_out/c28_ops_b64.csv | head -30

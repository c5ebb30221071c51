#!/bin/bash
# cluster-fused GroupNorm (finalize + apply in one launch at H*W <= 256): bit-identity test, U-Net tests, per-op and graph A/B
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "groupnorm or finalize" > gpurun_out/c34_gn_tests.log 2>&1; echo "gn tests rc=$?"; tail -5 gpurun_out/c34_gn_tests.log
timeout -k 5 400 python -m pytest tests/test_unet_gpu.py tests/test_search_gpu.py -x -q -p no:cacheprovider > gpurun_out/c34_unet_tests.log 2>&1; echo "unet tests rc=$?"; tail -3 gpurun_out/c34_unet_tests.log
for m in 1 0 1 0; do
  B200NS_GN_CLUSTER=$m timeout -k 5 300 python tools/profile_ops.py 64 --graph --csv gpurun_out/c34_ops_cl$m.csv > gpurun_out/c34_profile_ops_cl$m.log 2>&1
  echo "GN_CLUSTER=$m"; sed -n 1,10p gpurun_out/c34_profile_ops_cl$m.log
done

#!/bin/bash
mkdir -p gpurun_out
for bn in 0 128 64 256; do
  echo "== B200NS_XF_BN=$bn"
  B200NS_XF_BN=$bn timeout -k 5 600 python tools/profile_ops.py 64 --csv gpurun_out/c29_ops_xf$bn.csv > gpurun_out/c29_profile_ops_xf$bn.log 2>&1
  sed -n 1,4p gpurun_out/c29_profile_ops_xf$bn.log
  grep -E "enc.32x32_block0.norm2\+qkv|enc.16x16_block0.norm2\+qkv|enc.8x8_block0.norm2\+qkv" gpurun_out/c29_ops_xf$bn.csv
done

#!/bin/bash
# small-batch NFE profiles (strong scaling: 32 and 8 candidates per GPU), per op and as a graph
mkdir -p gpurun_out
for B in 32 8; do
  timeout -k 5 300 python tools/profile_ops.py $B --graph --csv gpurun_out/c41_ops_b$B.csv > gpurun_out/c41_profile_ops_b$B.log 2>&1
  sed -n 1,10p gpurun_out/c41_profile_ops_b$B.log
done

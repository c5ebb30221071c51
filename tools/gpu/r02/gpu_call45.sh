#!/bin/bash
# cluster-fused GroupNorm at small candidate batches (strong scaling): graph NFE with / without
mkdir -p gpurun_out
for B in 32 8; do for m in 0 1 0 1; do
  B200NS_GN_CLUSTER=$m timeout -k 5 200 python tools/profile_ops.py $B --graph > gpurun_out/c45_ops_b${B}_cl$m.log 2>&1
  echo "B=$B GN_CLUSTER=$m: $(sed -n 2p gpurun_out/c45_ops_b${B}_cl$m.log) | $(grep -E 'gn_norm|gn_finalize|gn_apply' gpurun_out/c45_ops_b${B}_cl$m.log | tr -s ' ' | tr '\n' ';')"
done; done

#!/bin/bash
mkdir -p gpurun_out
for pref in 0 128 64 256; do
  echo "== B200NS_PREC_BN_PREF=$pref"
  B200NS_PREC_BN_PREF=$pref timeout -k 5 600 python tools/profile_precise.py 1 2 4 > gpurun_out/c25_profile_precise_bn$pref.log 2>&1
  grep -o '"precise_R": [0-9]*, "nfe_ms_graph": [0-9.]*' gpurun_out/c25_profile_precise_bn$pref.log
  grep -o '"gemm_prec": [0-9.]*' gpurun_out/c25_profile_precise_bn$pref.log
done

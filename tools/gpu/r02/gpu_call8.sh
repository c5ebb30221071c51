#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "precise_pass/" --kernel-name regex:gemm_prec_kernel --launch-skip 55 --launch-count 4 -o gpurun_out/c8_prec_gemm_8x8 -f python tools/ncu_precise.py 2 > gpurun_out/c8_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/c8_prec_gemm_8x8.ncu-rep --page raw --csv > gpurun_out/c8_prec_gemm_8x8_raw.csv 2>/dev/null
ls -la gpurun_out/c8_*

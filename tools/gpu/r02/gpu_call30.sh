#!/bin/bash
mkdir -p gpurun_out
OPS="enc.32x32_block0.norm2+qkv"
B200NS_PDL=0 timeout -k 5 400 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/c30_prof_xf -f python tools/profile_one.py 64 $OPS > gpurun_out/c30_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/c30_prof_xf.ncu-rep --page source --csv > gpurun_out/c30_prof_xf_source.csv 2>/dev/null
ncu -i gpurun_out/c30_prof_xf.ncu-rep --page raw --csv > gpurun_out/c30_prof_xf_raw.csv 2>/dev/null
ls -la gpurun_out/c30_*

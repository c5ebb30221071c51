#!/bin/bash
mkdir -p gpurun_out
OPS="dec.64x64_up.conv0 dec.32x32_up.conv0 dec.64x64_up.conv1"
timeout -k 10 200 python tools/profile_one.py 64 $OPS > gpurun_out/up_one_plain.log 2>&1; echo "plain exit $?"; tail -n 4 gpurun_out/up_one_plain.log
timeout -k 10 250 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_up_one -f python tools/profile_one.py 64 $OPS > gpurun_out/up_one_ncu.log 2>&1; echo "ncu exit $?"; tail -n 2 gpurun_out/up_one_ncu.log
ncu -i gpurun_out/prof_up_one.ncu-rep --page raw --csv > gpurun_out/prof_up_one_raw.csv 2>/dev/null; ls -la gpurun_out/prof_up_one_raw.csv

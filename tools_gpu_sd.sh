#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_sd_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/test_sd_gpu.log 2>&1; echo "pytest exit $?"; tail -n 30 gpurun_out/test_sd_gpu.log
timeout -k 10 600 python tools/profile_sd.py 64 --beam 2 16 2 --csv gpurun_out/sd_ops_m64.csv > gpurun_out/sd_prof_m64.log 2>&1; echo "prof exit $?"; tail -n 60 gpurun_out/sd_prof_m64.log
timeout -k 10 600 python tools/profile_sd.py 128 > gpurun_out/sd_prof_m128.log 2>&1; echo "prof exit $?"; head -n 14 gpurun_out/sd_prof_m128.log

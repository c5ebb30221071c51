#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
timeout -k 10 600 python tools/profile_sd.py 64 --beam 2 16 2 --csv gpurun_out/sd_ops_m64_v3.csv > gpurun_out/sd_prof_m64_v3.log 2>&1; echo "prof exit $?"; head -n 30 gpurun_out/sd_prof_m64_v3.log; tail -n 1 gpurun_out/sd_prof_m64_v3.log

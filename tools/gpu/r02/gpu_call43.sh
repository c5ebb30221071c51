#!/bin/bash
mkdir -p gpurun_out
PYTHONHASHSEED=0 timeout -k 5 400 python tools/profile_eps04.py > gpurun_out/c43_profile_eps04.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c43_profile_eps04.log

#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python tools/profile_ops.py 64 --csv gpurun_out/c13_ops_b64.csv > gpurun_out/c13_profile_ops.log 2>&1; echo "profile rc=$?"
tail -50 gpurun_out/c13_profile_ops.log

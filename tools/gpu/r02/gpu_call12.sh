#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_clip_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/c12_clip_tests.log 2>&1; echo "clip tests rc=$?"
timeout -k 5 600 python tools/profile_clip.py 16 --csv gpurun_out/c12_clip_ops_b16.csv > gpurun_out/c12_profile_clip.log 2>&1; echo "profile rc=$?"
tail -30 gpurun_out/c12_clip_tests.log
tail -30 gpurun_out/c12_profile_clip.log

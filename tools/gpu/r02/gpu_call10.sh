#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c10_status.txt
timeout -k 5 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c10_gpu_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/c10_status.txt
timeout -k 5 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c10_profile_precise.log 2>&1; echo "profile rc=$?" >> gpurun_out/c10_status.txt
timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c10_bench_quick.json 2> gpurun_out/c10_bench_quick.err; echo "bench rc=$?" >> gpurun_out/c10_status.txt
cat gpurun_out/c10_status.txt
grep -E "passed|failed" gpurun_out/c10_gpu_suite.log | tail -2

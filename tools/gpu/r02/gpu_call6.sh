#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c6_status.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c6_gpu_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/c6_status.txt
timeout 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c6_profile_precise.log 2>&1; echo "profile rc=$?" >> gpurun_out/c6_status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "precise_pass/" --csv --log-file gpurun_out/c6_precise_launches_R2.csv python tools/ncu_precise.py 2 > gpurun_out/c6_ncu_precise.log 2>&1; echo "ncu precise rc=$?" >> gpurun_out/c6_status.txt
timeout 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c6_bench_quick.json 2> gpurun_out/c6_bench_quick.err; echo "bench rc=$?" >> gpurun_out/c6_status.txt
cat gpurun_out/c6_status.txt
grep -E "passed|failed" gpurun_out/c6_gpu_suite.log | tail -2

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c5_status.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c5_gpu_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/c5_status.txt
timeout 1500 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?" >> gpurun_out/c5_status.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/c5_status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "precise_pass/" --csv --log-file gpurun_out/c5_precise_launches_R2.csv python tools/ncu_precise.py 2 > gpurun_out/c5_ncu_precise.log 2>&1; echo "ncu precise rc=$?" >> gpurun_out/c5_status.txt
cat gpurun_out/c5_status.txt
grep -E "passed|failed" gpurun_out/c5_gpu_suite.log | tail -2
tail -c 600 gpurun_out/c5_bench.err

#!/bin/bash
# round-2 GPU call 1: precise-path unit tests, full-size parity measurement, regression suite, bench A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_precise_gpu.py -q -s -p no:cacheprovider > gpurun_out/c1_precise_tests.log 2>&1; echo "precise rc=$?" >> gpurun_out/c1_status.txt
timeout 900 python -m pytest tests/test_full_parity_gpu.py -q -s -p no:cacheprovider > gpurun_out/c1_full_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/c1_status.txt
timeout 600 python tools/profile_precise.py 1 2 4 8 > gpurun_out/c1_profile_precise.log 2>&1; echo "profile rc=$?" >> gpurun_out/c1_status.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --ignore=tests/test_precise_gpu.py --ignore=tests/test_full_parity_gpu.py > gpurun_out/c1_gpu_suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/c1_status.txt
timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 0 > gpurun_out/c1_bench_esc0.json 2> gpurun_out/c1_bench_esc0.err; echo "bench0 rc=$?" >> gpurun_out/c1_status.txt
timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 1 > gpurun_out/c1_bench_esc1.json 2> gpurun_out/c1_bench_esc1.err; echo "bench1 rc=$?" >> gpurun_out/c1_status.txt
cat gpurun_out/c1_status.txt
tail -5 gpurun_out/c1_precise_tests.log
tail -5 gpurun_out/c1_full_parity.log

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c2_status.txt
timeout 900 python -m pytest tests/test_precise_gpu.py tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_search_gpu.py tests/test_full_size_gpu.py tests/test_classifier_gpu.py -q -s -p no:cacheprovider > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c2_status.txt
B200NS_ANALYSIS=1 timeout 900 python -m pytest tests/test_full_parity_gpu.py -q -s -p no:cacheprovider -k analysis > gpurun_out/c2_analysis_precise.log 2>&1; echo "analysis rc=$?" >> gpurun_out/c2_status.txt
B200NS_ANALYSIS=1 B200NS_PREC_NOLO=1 timeout 900 python -m pytest tests/test_full_parity_gpu.py -q -s -p no:cacheprovider -k analysis > gpurun_out/c2_analysis_nolo.log 2>&1; echo "analysis nolo rc=$?" >> gpurun_out/c2_status.txt
timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 1 > gpurun_out/c2_bench_esc1.json 2> gpurun_out/c2_bench_esc1.err; echo "bench1 rc=$?" >> gpurun_out/c2_status.txt
cat gpurun_out/c2_status.txt
grep -E "passed|failed" gpurun_out/c2_tests.log | tail -3

#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c7_status.txt
timeout 1200 python -m pytest tests/test_precise_gpu.py tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_search_gpu.py -q -p no:cacheprovider > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c7_status.txt
timeout 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c7_profile_precise.log 2>&1; echo "profile rc=$?" >> gpurun_out/c7_status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "precise_pass/" --csv --log-file gpurun_out/c7_precise_launches_R2.csv python tools/ncu_precise.py 2 > gpurun_out/c7_ncu_precise.log 2>&1; echo "ncu precise rc=$?" >> gpurun_out/c7_status.txt
timeout 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c7_bench_quick.json 2> gpurun_out/c7_bench_quick.err; echo "bench rc=$?" >> gpurun_out/c7_status.txt
B200NS_LANES=2 B200NS_LANES_SEQ=1 timeout 900 python bench.py --quick --no-cpu-baseline --escalate 0 > gpurun_out/c7_bench_lanes2seq.json 2> gpurun_out/c7_bench_lanes2seq.err; echo "bench l2 rc=$?" >> gpurun_out/c7_status.txt
B200NS_LANES=4 B200NS_LANES_SEQ=1 timeout 900 python bench.py --quick --no-cpu-baseline --escalate 0 > gpurun_out/c7_bench_lanes4seq.json 2> gpurun_out/c7_bench_lanes4seq.err; echo "bench l4 rc=$?" >> gpurun_out/c7_status.txt
cat gpurun_out/c7_status.txt
grep -E "passed|failed" gpurun_out/c7_tests.log | tail -2

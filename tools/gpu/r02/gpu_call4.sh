#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c4_status.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/c4_gpu_suite_fp16.log 2>&1; echo "suite fp16 rc=$?" >> gpurun_out/c4_status.txt
cp gpurun_out/parity_full_escalated_tables.pt gpurun_out/c4_fp16_escalated_tables.pt; cp gpurun_out/parity_full_eps04_tables.pt gpurun_out/c4_fp16_eps04_tables.pt; cp gpurun_out/parity_full_bf16_tables.pt gpurun_out/c4_fp16_plain_tables.pt
timeout 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c4_profile_precise.log 2>&1; echo "profile rc=$?" >> gpurun_out/c4_status.txt
for k in 0.3 0.5; do
timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 1 --kappa $k > gpurun_out/c4_bench_fp16_esc1_k$k.json 2> gpurun_out/c4_bench_fp16_esc1_k$k.err; echo "bench fp16 esc1 k$k rc=$?" >> gpurun_out/c4_status.txt
done
timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 0 > gpurun_out/c4_bench_fp16_esc0.json 2> gpurun_out/c4_bench_fp16_esc0.err; echo "bench fp16 esc0 rc=$?" >> gpurun_out/c4_status.txt
B200NS_ACT=bf16 timeout 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline --escalate 0 > gpurun_out/c4_bench_bf16_esc0.json 2> gpurun_out/c4_bench_bf16_esc0.err; echo "bench bf16 esc0 rc=$?" >> gpurun_out/c4_status.txt
B200NS_ACT=bf16 timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/c4_gpu_suite_bf16.log 2>&1; echo "suite bf16 rc=$?" >> gpurun_out/c4_status.txt
cat gpurun_out/c4_status.txt
grep -E "passed|failed" gpurun_out/c4_gpu_suite_fp16.log gpurun_out/c4_gpu_suite_bf16.log | tail -3

#!/bin/bash
# finalize kernel with 8 loads in flight (order-preserving): kernel + parity + sharded tests, per-op profile, CLI runs, quick bench
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_full_parity_gpu.py tests/test_search_gpu.py tests/test_sharded_gpu.py -x -q -p no:cacheprovider > gpurun_out/c40_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c40_tests.log
timeout -k 5 300 python tools/profile_ops.py 64 --graph --csv gpurun_out/c40_ops.csv > gpurun_out/c40_profile_ops.log 2>&1; sed -n 1,10p gpurun_out/c40_profile_ops.log
( time PYTHONHASHSEED=0 timeout -k 5 300 python main.py --backend edm --scorer brightness --method eps_greedy --N 16 --K 2 --output gpurun_out/c40_cli_eps_greedy.png ) > gpurun_out/c40_cli_eps_greedy.log 2>&1; echo "cli eps_greedy rc=$?"; tail -4 gpurun_out/c40_cli_eps_greedy.log
( time timeout -k 5 300 python main.py --backend edm --scorer compressibility --method zero_order --N 8 --K 1 --output gpurun_out/c40_cli_zero_order.png ) > gpurun_out/c40_cli_zero_order.log 2>&1; echo "cli zero_order rc=$?"; tail -4 gpurun_out/c40_cli_zero_order.log
timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c40_bench_quick.json 2> gpurun_out/c40_bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c40_bench_quick.json') if l.startswith('{')][-1])
x=d['extras']
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(x['no_escalation']['ms_per_step'],2), d['roofline']['ms_by_kernel_kind'], d['clocks'])
P

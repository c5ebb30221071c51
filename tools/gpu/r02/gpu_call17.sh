#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_precise_gpu.py tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_kernels_gpu.py tests/test_search_gpu.py -x -q -p no:cacheprovider > gpurun_out/c17_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c17_tests.log
timeout -k 5 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c17_profile_precise_pdl.log 2>&1; echo "profile rc=$?"; grep -o '"precise_R": [0-9]*, "nfe_ms_graph": [0-9.]*' gpurun_out/c17_profile_precise_pdl.log
B200NS_PREC_PDL=0 timeout -k 5 600 python tools/profile_precise.py 1 2 4 > gpurun_out/c17_profile_precise_nopdl.log 2>&1; grep -o '"precise_R": [0-9]*, "nfe_ms_graph": [0-9.]*' gpurun_out/c17_profile_precise_nopdl.log
timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c17_bench_quick.json 2> gpurun_out/c17_bench_quick.err; echo "bench rc=$?"
B200NS_PREC_PDL=0 timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c17_bench_quick_nopdl.json 2> gpurun_out/c17_bench_quick_nopdl.err; echo "bench nopdl rc=$?"
python - <<'P'
import json
for f in ('c17_bench_quick','c17_bench_quick_nopdl'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],2), 'no_esc', round(d['extras']['no_escalation']['value'],1))
P

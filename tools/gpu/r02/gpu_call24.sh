#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_precise_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/c24_precise_tests.log 2>&1; echo "precise tests rc=$?"; grep -E "rel |passed|failed|Error" gpurun_out/c24_precise_tests.log | tail -12
timeout -k 5 600 python tools/profile_precise.py 1 2 3 4 8 > gpurun_out/c24_profile_precise.log 2>&1; echo "profile rc=$?"; grep -o '"precise_R": [0-9]*, "nfe_ms_graph": [0-9.]*' gpurun_out/c24_profile_precise.log
grep -o '"by_kind": {[^}]*}' gpurun_out/c24_profile_precise.log | head -2
timeout -k 5 900 python -m pytest tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_search_gpu.py -x -q -p no:cacheprovider > gpurun_out/c24_parity_tests.log 2>&1; echo "parity tests rc=$?"; tail -3 gpurun_out/c24_parity_tests.log
timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c24_bench_quick.json 2> gpurun_out/c24_bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c24_bench_quick.json') if l.startswith('{')][-1])
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['ms_per_step'],2), d['clocks'], d['escalation']['rows_refined_per_step'])
P

#!/bin/bash
# noise inputs prepared one round ahead of the escalation sync: search / parity / sharded / full-size tests, bench with the eps = 0.4 extra
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_search_gpu.py tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_full_size_gpu.py -x -q -p no:cacheprovider > gpurun_out/c42_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c42_tests.log
timeout -k 5 900 python bench.py --no-cpu-baseline > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c42_bench.json') if l.startswith('{')][-1])
x=d['extras']
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'no_esc', round(x['no_escalation']['ms_per_step'],2), 'eps04', round(x['eps04']['ms_per_step'],2), 'spec', round(x['speculation']['ms_per_step'],2), 'recompute', round(x['commit_recompute']['ms_per_step'],2), 'strong3', round(x['strong_config3']['ms_per_step'],1), d['clocks'])
P

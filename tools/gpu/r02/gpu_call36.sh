#!/bin/bash
# speculation gated by the 16-bit winner's lead: tests, then bench with two gate values (the log shows the leads of the missed rounds)
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_search_gpu.py tests/test_sharded_gpu.py -x -q -p no:cacheprovider > gpurun_out/c36_search_tests.log 2>&1; echo "search tests rc=$?"; tail -5 gpurun_out/c36_search_tests.log
for gap in 0.08 0.16; do
B200NS_SPEC_GAP=$gap timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c36_bench_quick_gap$gap.json 2> gpurun_out/c36_bench_quick_gap$gap.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/c36_bench_quick_gap$gap.json') if l.startswith('{')][-1])
x=d['extras']
e=d['escalation']
print('gap $gap bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_spec', round(x['no_speculation']['ms_per_step'],2), 'no_esc', round(x['no_escalation']['ms_per_step'],2), 'missed', e['missed_steps'], e['log_step_lead_speculated'], d['clocks'])
P
done

#!/bin/bash
# speculation past escalated rounds: bit-identity tests (incl. forced rollbacks), parity + sharded suites, bench A/B
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_search_gpu.py -x -q -p no:cacheprovider > gpurun_out/c35_search_tests.log 2>&1; echo "search tests rc=$?"; tail -5 gpurun_out/c35_search_tests.log
timeout -k 5 600 python -m pytest tests/test_full_parity_gpu.py tests/test_sharded_gpu.py tests/test_full_size_gpu.py -x -q -p no:cacheprovider > gpurun_out/c35_parity_tests.log 2>&1; echo "parity tests rc=$?"; tail -5 gpurun_out/c35_parity_tests.log
for prio in -1 0; do
B200NS_SPEC_PRIO=$prio timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c35_bench_quick_prio$prio.json 2> gpurun_out/c35_bench_quick_prio$prio.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/c35_bench_quick_prio$prio.json') if l.startswith('{')][-1])
x=d['extras']
print('prio $prio bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_spec', round(x['no_speculation']['ms_per_step'],2), 'no_esc', round(x['no_escalation']['ms_per_step'],2), d['escalation'], d['clocks'])
P
done

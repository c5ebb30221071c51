#!/bin/bash
# final code of the round, the driver's own sequence: GPU suite, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c37_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/c37_gpu_suite.log
timeout -k 5 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c37_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c37_smoke.log
( time timeout -k 5 1500 python bench.py > gpurun_out/c37_bench_1gpu.json 2> gpurun_out/c37_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout -k 5 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c37_bench_reference_arm.json 2> gpurun_out/c37_bench_ref.err ) 2>&1 | grep real; echo "ref rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c37_bench_1gpu.json') if l.startswith('{')][-1])
x=d['extras']
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])
print({k:(round(v['value'],1) if isinstance(v,dict) and 'value' in v else None) for k,v in x.items()})
r=json.loads([l for l in open('gpurun_out/c37_bench_reference_arm.json') if l.startswith('{')][-1]); print('ref', r['value'], r.get('cpu_baseline'))
P

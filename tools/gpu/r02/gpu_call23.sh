#!/bin/bash
mkdir -p gpurun_out
for k in 1 2; do
  timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c23_bench_$k.json 2> gpurun_out/c23_bench_$k.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/c23_bench_$k.json') if l.startswith('{')][-1])
print('run $k:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'no_esc', round(d['extras']['no_escalation']['ms_per_step'],2), d['clocks'])
P
done

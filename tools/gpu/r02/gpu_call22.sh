#!/bin/bash
mkdir -p gpurun_out
for ms in 100 200 1000 100 1000; do
  B200NS_CLOCK_MS=$ms timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c22_bench_clk$ms.json 2> gpurun_out/c22_bench_clk$ms.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/c22_bench_clk$ms.json') if l.startswith('{')][-1])
print('clock poll $ms ms:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'no_esc', round(d['extras']['no_escalation']['ms_per_step'],2), d['clocks'])
P
done

#!/bin/bash
mkdir -p gpurun_out
for P in 0 2 0 2; do
B200NS_PDL=$P timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl_$P.json 2> gpurun_out/bench_pdl_$P.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl_$P.json').read().strip().splitlines()[-1]);print('PDL=$P', round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),d['clocks']['sm_mhz'])"
done

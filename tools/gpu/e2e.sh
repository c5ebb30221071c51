#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 1" "1 1" "0 1" "1 1" "0 1" "0 0"; do
set -- $cfg
timeout 600 python bench.py --no-cpu-baseline --prefetch $1 --async-readback $2 > gpurun_out/bench_e2e_$1$2.json 2> gpurun_out/bench_e2e_$1$2.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_e2e_$1$2.json').read().strip().splitlines()[-1]);print('prefetch=$1 async=$2', round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],3))"
done

#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest_gpu.log
B200NS_FUSED_UP=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_noup.json 2> gpurun_out/bench_noup.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_noup.json').read().strip().splitlines()[-1]);print('FUSED_UP=0', round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),d['roofline']['ms_by_kernel_kind'])"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_up.json 2> gpurun_out/bench_up.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_up.json').read().strip().splitlines()[-1]);print('FUSED_UP=1', round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),d['roofline']['ms_by_kernel_kind'])"; tail -n 3 gpurun_out/bench_up.err

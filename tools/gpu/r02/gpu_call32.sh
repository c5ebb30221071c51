#!/bin/bash
# final code: full suite (fp16 default), full suite with the bfloat16-storage build, quick bench
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c32_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/c32_gpu_suite.log
B200NS_ACT=bf16 timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c32_gpu_suite_bf16.log 2>&1; echo "suite bf16 rc=$?"; tail -8 gpurun_out/c32_gpu_suite_bf16.log
timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c32_bench_quick.json 2> gpurun_out/c32_bench_quick.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c32_bench_quick.json') if l.startswith('{')][-1])
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['ms_per_step'],2), d['clocks'])
P

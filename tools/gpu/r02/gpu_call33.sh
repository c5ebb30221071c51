#!/bin/bash
mkdir -p gpurun_out
B200NS_ACT=bf16 timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c33_gpu_suite_bf16.log 2>&1; echo "suite bf16 rc=$?"; tail -4 gpurun_out/c33_gpu_suite_bf16.log
B200NS_ACT=bf16 timeout -k 5 600 python bench.py --quick --no-cpu-baseline > gpurun_out/c33_bench_quick_bf16.json 2> gpurun_out/c33_bench_quick_bf16.err; echo "bench bf16 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c33_bench_quick_bf16.json') if l.startswith('{')][-1])
print('bench bf16:', round(d['value'],1), round(d['ms_per_step'],2), 'no_esc', round(d['extras']['no_escalation']['value'],1), d['dtype'], d['escalation']['rows_refined_per_step'])
P

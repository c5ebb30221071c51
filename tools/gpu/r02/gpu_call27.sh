#!/bin/bash
mkdir -p gpurun_out
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout -k 10 900 $TR --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${G}gpu.json 2> gpurun_out/c27_bench_${G}gpu.err
echo "${G}gpu bench exit $?"; python - <<P
import json
for l in open('gpurun_out/r02_bench_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$G GPUs', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['value'],1), 'strong3', d['extras'].get('strong_config3',{}).get('value'), d['escalation']['rows_refined_per_step'])
P
tail -n 3 gpurun_out/c27_bench_${G}gpu.err
timeout -k 10 600 $TR --master-port 29513 tools/bench_sd_beam.py --gpus $G --steps 6 --warmup 2 > gpurun_out/r02_bench_sd_beam_${G}gpu.json 2> gpurun_out/c27_sd_beam_${G}gpu.err
echo "sd beam exit $?"; python - <<P
import json
for l in open('gpurun_out/r02_bench_sd_beam_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('SD beam $G GPUs', round(d['value'],1), round(d['ms_per_step'],2))
P

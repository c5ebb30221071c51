#!/bin/bash
# final code on 8 GPUs over NCCL: sharded == unsharded (N=256), bench (weak + strong extras)
mkdir -p gpurun_out
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29512 tools/check_sharded.py --N 256 --steps 6 > gpurun_out/c39_sharded_N256_${G}gpu.json 2> gpurun_out/c39_sharded_N256_${G}gpu.err; echo "sharded N256 exit $?"; tail -n 1 gpurun_out/c39_sharded_N256_${G}gpu.json | cut -c1-420
timeout -k 10 600 $TR --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/c39_bench_${G}gpu.json 2> gpurun_out/c39_bench_${G}gpu.err
echo "${G}gpu bench exit $?"; python - <<P
import json
for l in open('gpurun_out/c39_bench_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$G GPUs', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['value'],1), 'eps04', round(d['extras']['eps04']['ms_per_step'],1), 'strong3', d['extras'].get('strong_config3',{}).get('value'), d['extras'].get('strong_config3',{}).get('ms_per_step'), d['escalation']['rows_refined_per_step'])
P
tail -n 2 gpurun_out/c39_bench_${G}gpu.err

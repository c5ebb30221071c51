#!/bin/bash
# multi-GPU evidence: sharded == unsharded (NCCL, escalation on) and bench (weak + strong_config3) on G GPUs
mkdir -p gpurun_out
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
PYTHONHASHSEED=0 timeout -k 10 400 $TR --master-port 29512 tools/check_sharded.py --N 256 --steps 4 > gpurun_out/r02_sharded_zero_order_N256_${G}gpu.json 2> gpurun_out/c20_sharded_${G}gpu.err
echo "sharded exit $?"; grep '"check"' gpurun_out/r02_sharded_zero_order_N256_${G}gpu.json | cut -c1-500
PYTHONHASHSEED=0 timeout -k 10 400 $TR --master-port 29514 tools/check_sharded.py --N 64 --K 2 --steps 4 > gpurun_out/r02_sharded_zero_order_N64_K2_${G}gpu.json 2> gpurun_out/c20_sharded_k2_${G}gpu.err
echo "sharded K2 exit $?"; grep '"check"' gpurun_out/r02_sharded_zero_order_N64_K2_${G}gpu.json | cut -c1-500
timeout -k 10 900 $TR --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${G}gpu.json 2> gpurun_out/c20_bench_${G}gpu.err
echo "${G}gpu bench exit $?"; python - <<P
import json
for l in open('gpurun_out/r02_bench_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$G GPUs', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['value'],1), 'strong3', d['extras'].get('strong_config3',{}).get('value'), d['escalation']['rows_refined_per_step'])
P
tail -n 3 gpurun_out/c20_bench_${G}gpu.err

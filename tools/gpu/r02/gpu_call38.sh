#!/bin/bash
# final code on G GPUs over NCCL: sharded == unsharded (N=256 K=1, N=64 K=2; also with speculation on), bench (weak + strong extras), reference arm under torchrun
mkdir -p gpurun_out
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29512 tools/check_sharded.py --N 256 --steps 6 > gpurun_out/c38_sharded_N256_${G}gpu.json 2> gpurun_out/c38_sharded_N256_${G}gpu.err; echo "sharded N256 exit $?"; tail -n 1 gpurun_out/c38_sharded_N256_${G}gpu.json | cut -c1-400
B200NS_SPECULATE=1 B200NS_SPEC_GAP=0 timeout -k 10 400 $TR --master-port 29514 tools/check_sharded.py --N 64 --K 2 --steps 6 > gpurun_out/c38_sharded_N64_K2_spec_${G}gpu.json 2> gpurun_out/c38_sharded_N64_K2_spec_${G}gpu.err; echo "sharded N64 K2 (speculation on) exit $?"; tail -n 1 gpurun_out/c38_sharded_N64_K2_spec_${G}gpu.json | cut -c1-400
timeout -k 10 900 $TR --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/c38_bench_${G}gpu.json 2> gpurun_out/c38_bench_${G}gpu.err
echo "${G}gpu bench exit $?"; python - <<P
import json
for l in open('gpurun_out/c38_bench_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$G GPUs', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['value'],1), 'strong3', d['extras'].get('strong_config3',{}).get('value'), d['escalation']['rows_refined_per_step'])
P
tail -n 3 gpurun_out/c38_bench_${G}gpu.err
timeout -k 10 600 $TR --master-port 29515 bench.py --impl reference --gpus $G --steps 2 --warmup 1 > gpurun_out/c38_bench_ref_${G}gpu.json 2> gpurun_out/c38_bench_ref_${G}gpu.err; echo "ref arm exit $?"; tail -n 1 gpurun_out/c38_bench_ref_${G}gpu.json | cut -c1-300

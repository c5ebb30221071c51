#!/bin/bash
# final code on 2 GPUs: the driver's torchrun bench (own arm + reference arm), sharded check; bf16-storage build: search / parity / sharded tests
mkdir -p gpurun_out
G=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29512 tools/check_sharded.py --N 64 --K 2 --steps 6 > gpurun_out/c44_sharded_N64_K2_${G}gpu.json 2> gpurun_out/c44_sharded_${G}gpu.err; echo "sharded N64 K2 exit $?"; tail -n 1 gpurun_out/c44_sharded_N64_K2_${G}gpu.json | cut -c1-330
timeout -k 10 600 $TR --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 > gpurun_out/c44_bench_${G}gpu.json 2> gpurun_out/c44_bench_${G}gpu.err
echo "${G}gpu bench exit $?"; python - <<P
import json
for l in open('gpurun_out/c44_bench_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$G GPUs', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'no_esc', round(d['extras']['no_escalation']['value'],1), 'eps04', round(d['extras']['eps04']['ms_per_step'],1), 'strong3', d['extras'].get('strong_config3',{}).get('value'), d['escalation']['rows_refined_per_step'], d['cpu_baseline'])
P
B200NS_ACT=bf16 timeout -k 5 600 python -m pytest tests/test_search_gpu.py tests/test_full_parity_gpu.py tests/test_sharded_gpu.py -x -q -p no:cacheprovider > gpurun_out/c44_tests_bf16.log 2>&1; echo "bf16 tests rc=$?"; tail -3 gpurun_out/c44_tests_bf16.log

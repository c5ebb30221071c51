#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py --N 256 --steps 2 > gpurun_out/sharded_2gpu.json 2> gpurun_out/sharded_2gpu.err
echo "sharded exit $?"; grep '"check"' gpurun_out/sharded_2gpu.json | cut -c1-400
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 18 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu bench exit $?"; python -c "
import json
for l in open('gpurun_out/bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('2 GPUs', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"

#!/bin/bash
mkdir -p gpurun_out
G=8
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 tools/bench_sd_beam.py --gpus $G --steps 6 --warmup 2 > gpurun_out/bench_sd_beam_${G}gpu.json 2> gpurun_out/bench_sd_beam_${G}gpu.err
echo "sd beam exit $?"; python -c "
import json
for l in open('gpurun_out/bench_sd_beam_${G}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('sd beam ${G} GPUs', d['value'], d['ms_per_step'], d['config']['workload'][:100])"
tail -n 2 gpurun_out/bench_sd_beam_${G}gpu.err

#!/bin/bash
# usage: tools_gpu_multi.sh G  -- sharded parity (config 3) + scaling bench at G GPUs
G=${1:-2}
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py --N 256 --steps 4 > gpurun_out/sharded_${G}gpu.json 2> gpurun_out/sharded_${G}gpu.err
echo "sharded check exit $?"; tail -n 2 gpurun_out/sharded_${G}gpu.json | cut -c1-700; tail -n 3 gpurun_out/sharded_${G}gpu.err
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err
echo "bench exit $?"; tail -n 1 gpurun_out/bench_${G}gpu.json | cut -c1-330; tail -n 3 gpurun_out/bench_${G}gpu.err

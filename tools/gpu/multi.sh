#!/bin/bash
# multi-GPU: bench (weak scaling), SD beam bench
mkdir -p gpurun_out
G=${1:-8}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 18 --warmup 3 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err
echo "${G}gpu bench exit $?"; tail -c 2600 gpurun_out/bench_${G}gpu.json; tail -n 4 gpurun_out/bench_${G}gpu.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 tools/bench_sd_beam.py --gpus $G --steps 6 --warmup 2 > gpurun_out/bench_sd_beam_${G}gpu.json 2> gpurun_out/bench_sd_beam_${G}gpu.err
echo "sd beam exit $?"; tail -c 2000 gpurun_out/bench_sd_beam_${G}gpu.json; tail -n 4 gpurun_out/bench_sd_beam_${G}gpu.err

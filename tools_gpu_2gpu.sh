#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 18 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu bench exit $?"; tail -c 2500 gpurun_out/bench_2gpu.json; tail -n 8 gpurun_out/bench_2gpu.err
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?"; cat gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err

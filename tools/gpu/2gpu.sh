#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py --N 256 --steps 3 > gpurun_out/sharded_2gpu.json 2> gpurun_out/sharded_2gpu.err
echo "sharded exit $?"; tail -n 3 gpurun_out/sharded_2gpu.json; tail -n 5 gpurun_out/sharded_2gpu.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_sharded.py --N 64 --steps 2 --K 2 > gpurun_out/sharded_2gpu_k2.json 2> gpurun_out/sharded_2gpu_k2.err
echo "sharded K2 exit $?"; tail -n 3 gpurun_out/sharded_2gpu_k2.json; tail -n 5 gpurun_out/sharded_2gpu_k2.err
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 18 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu bench exit $?"; tail -c 2500 gpurun_out/bench_2gpu.json; tail -n 8 gpurun_out/bench_2gpu.err
timeout -k 10 900 python bench.py --steps 18 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 2600 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err

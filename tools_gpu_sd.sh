#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest_gpu.log
timeout -k 10 600 python main.py --backend sd --scorer brightness --method beam --B 2 --N 4 --steps 3 --output gpurun_out/sd_beam.png > gpurun_out/cli_sd_beam.log 2>&1; echo "cli beam exit $?"; tail -n 4 gpurun_out/cli_sd_beam.log
timeout -k 10 600 python main.py --backend sd --scorer brightness --method eps_greedy --N 4 --K 2 --steps 3 --output gpurun_out/sd_eg.png > gpurun_out/cli_sd_eg.log 2>&1; echo "cli eps_greedy exit $?"; tail -n 4 gpurun_out/cli_sd_eg.log
timeout -k 10 600 python main.py --backend edm --scorer brightness --method eps_greedy --N 8 --K 1 --output gpurun_out/edm_eg.png > gpurun_out/cli_edm_eg.log 2>&1; echo "cli edm exit $?"; tail -n 4 gpurun_out/cli_edm_eg.log

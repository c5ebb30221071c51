#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_conv_kernel -s 60 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -12

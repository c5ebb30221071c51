#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one full capture of the dominant kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline"
timeout -k 5 600 $CMD > gpurun_out/c14_plain.json 2> gpurun_out/c14_plain.err; echo "plain rc=$?"
timeout -k 5 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 15000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/c14_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:gemm_conv_kernel -s 400 -c 8 -o gpurun_out/r02_prof_gemm -f $CMD > gpurun_out/c14_ncu_full.log 2>&1; echo "ncu full gemm rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:"gn_apply_kernel|attention_kernel_v3" -s 100 -c 8 -o gpurun_out/r02_prof_gn_attn -f $CMD > gpurun_out/c14_ncu_full2.log 2>&1; echo "ncu full gn/attn rc=$?"
ncu -i gpurun_out/r02_prof_gemm.ncu-rep --page raw --csv > gpurun_out/r02_prof_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_gn_attn.ncu-rep --page raw --csv > gpurun_out/r02_prof_gn_attn_raw.csv 2>/dev/null
ls -la gpurun_out | grep -E "r02_|c14_"

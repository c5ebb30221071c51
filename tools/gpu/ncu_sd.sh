#!/bin/bash
mkdir -p gpurun_out
OPS="up_blocks.3.attentions.1.transformer_blocks.0.attn1.attn up_blocks.3.attentions.1.transformer_blocks.0.attn2.xattn up_blocks.1.attentions.1.transformer_blocks.0.attn1.attn up_blocks.3.attentions.1.transformer_blocks.0.norm1 up_blocks.3.attentions.1.transformer_blocks.0.ff.proj up_blocks.3.attentions.1.transformer_blocks.0.ff.geglu up_blocks.3.resnets.0.conv1[192] up_blocks.3.resnets.0.conv1[128]"
timeout -k 10 600 python tools/profile_one.py 64 --sd $OPS > gpurun_out/sd_one_plain.log 2>&1; echo "plain exit $?"; tail -n 9 gpurun_out/sd_one_plain.log
timeout -k 10 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_sd_one -f python tools/profile_one.py 64 --sd $OPS > gpurun_out/sd_one_ncu.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/sd_one_ncu.log
ncu -i gpurun_out/prof_sd_one.ncu-rep --page raw --csv > gpurun_out/prof_sd_one_raw.csv 2>/dev/null; ls -la gpurun_out/prof_sd_one*

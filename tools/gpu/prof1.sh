#!/bin/bash
mkdir -p gpurun_out
OPS="${OPS:-enc.32x32_block0.qkv enc.32x32_block0.proj enc.64x64_block0.conv1 dec.64x64_block0.norm0.finalize dec.64x64_block0.norm0.apply enc.32x32_block0.attn}"
python tools/profile_one.py 64 $OPS > gpurun_out/one_plain.log 2>&1; echo "plain exit $?"; tail -n 8 gpurun_out/one_plain.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_one -f python tools/profile_one.py 64 $OPS > gpurun_out/one_ncu.log 2>&1; echo "ncu exit $?"; tail -n 3 gpurun_out/one_ncu.log

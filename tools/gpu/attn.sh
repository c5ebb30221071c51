#!/bin/bash
mkdir -p gpurun_out
B200NS_ATTN_ALIAS=1 timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py tests/test_sd_gpu.py tests/test_unet_gpu.py -q -m gpu -p no:cacheprovider -x -k "attention or unet or adm" > gpurun_out/pytest_attn.log 2>&1; echo "pytest(alias) exit $?"; tail -n 5 gpurun_out/pytest_attn.log
for P in 0 1 0 1; do
echo "== B200NS_ATTN_ALIAS=$P"
B200NS_ATTN_ALIAS=$P timeout -k 10 300 python tools/profile_one.py 64 --sd up_blocks.3.attentions.1.transformer_blocks.0.attn1.attn 2>&1 | grep "TFLOP"
B200NS_ATTN_ALIAS=$P timeout -k 10 300 python tools/profile_one.py 64 enc.32x32_block0.attn enc.16x16_block0.attn 2>&1 | grep "TFLOP"
done

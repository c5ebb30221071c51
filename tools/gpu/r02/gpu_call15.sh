#!/bin/bash
# round-2 ncu evidence, second attempt: PDL off (ncu cannot profile programmatically-serialised graph nodes), bounded launch count
mkdir -p gpurun_out
export B200NS_PDL=0
CMD="python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline"
timeout -k 5 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/c15_ncu_launch.log 2>&1; echo "launch list rc=$?"
OPS="dec.64x64_up.conv1 dec.32x32_block0.conv0 enc.64x64_block0.conv1 enc.32x32_block0.qkv enc.32x32_block0.proj dec.64x64_block0.norm0.finalize dec.64x64_block0.norm0.apply enc.32x32_block0.attn"
timeout -k 5 300 python tools/profile_one.py 64 $OPS > gpurun_out/c15_one_plain.log 2>&1; echo "plain rc=$?"; tail -n 9 gpurun_out/c15_one_plain.log
timeout -k 5 400 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_prof_one -f python tools/profile_one.py 64 $OPS > gpurun_out/c15_one_ncu.log 2>&1; echo "ncu one rc=$?"
ncu -i gpurun_out/r02_prof_one.ncu-rep --page raw --csv > gpurun_out/r02_prof_one_raw.csv 2>/dev/null
unset B200NS_PDL
timeout -k 5 600 python tools/bench_sd_beam.py --steps 4 --warmup 2 --clip > gpurun_out/c15_sd_beam_clip.json 2> gpurun_out/c15_sd_beam_clip.err; echo "sd beam clip rc=$?"
timeout -k 5 600 python tools/bench_sd_beam.py --steps 4 --warmup 2 --vae > gpurun_out/c15_sd_beam_vae.json 2> gpurun_out/c15_sd_beam_vae.err; echo "sd beam vae rc=$?"
ls -la gpurun_out | grep -E "r02_|c15_"

#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/profile_sd.py 64 --csv gpurun_out/sd_ops_m64_v4.csv > gpurun_out/sd_prof_m64_v4.log 2>&1; echo "prof exit $?"; head -n 12 gpurun_out/sd_prof_m64_v4.log
timeout -k 10 600 python tools/profile_vae.py 8 > gpurun_out/vae_prof_b8.log 2>&1; head -n 9 gpurun_out/vae_prof_b8.log
timeout 600 python tools/bench_sd_beam.py --steps 6 --warmup 2 > gpurun_out/bench_sd_beam_1gpu.json 2> gpurun_out/bench_sd_beam_1gpu.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_sd_beam_1gpu.json').read().strip().splitlines()[-1]);print('sd beam', d['value'], d['ms_per_step'])"
timeout 600 python tools/bench_sd_beam.py --steps 4 --warmup 2 --vae > gpurun_out/bench_sd_beam_vae_1gpu.json 2> gpurun_out/bench_sd_beam_vae_1gpu.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_sd_beam_vae_1gpu.json').read().strip().splitlines()[-1]);print('sd beam+vae', d['value'], d['ms_per_step'])"

#!/bin/bash
# round-2 ncu launch list of the TIMED REGION of the bench command (cudaProfilerStart/Stop window), PDL off for ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline"
B200NS_PDL=0 timeout -k 5 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/c16_ncu_launch.log 2>&1; echo "launch list rc=$?"
tail -2 gpurun_out/c16_ncu_launch.log
timeout -k 5 600 python tools/bench_sd_beam.py --steps 4 --warmup 2 --clip > gpurun_out/c16_sd_beam_clip.json 2> gpurun_out/c16_sd_beam_clip.err; echo "sd beam clip rc=$?"
timeout -k 5 600 python tools/bench_sd_beam.py --steps 4 --warmup 2 --vae > gpurun_out/c16_sd_beam_vae.json 2> gpurun_out/c16_sd_beam_vae.err; echo "sd beam vae rc=$?"
timeout -k 5 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c16_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/c16_gpu_suite.log
timeout -k 5 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c16_smoke.log

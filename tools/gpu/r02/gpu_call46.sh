#!/bin/bash
# end of round 2: the driver's sequence on the final commit (GPU suite, smoke, default bench, reference arm), then the ncu launch list
# of the timed region of the bench command (after the plain run has exited 0)
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c46_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/c46_gpu_suite.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c46_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c46_smoke.log
timeout -k 5 900 python bench.py > gpurun_out/c46_bench_1gpu.json 2> gpurun_out/c46_bench.err; echo "bench rc=$?"
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c46_bench_reference_arm.json 2> gpurun_out/c46_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/c46_bench_1gpu.json') if l.startswith('{')][-1])
x=d['extras']
print('bench:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), d['clocks'])
print({k:(round(v['value'],1) if isinstance(v,dict) and 'value' in v else None) for k,v in x.items()})
P
CMD="python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline"
timeout -k 5 300 $CMD > gpurun_out/c46_plain_run.json 2> gpurun_out/c46_plain_run.err; echo "plain rc=$?"
B200NS_PDL=0 timeout -k 5 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/c46_launches.csv $CMD > gpurun_out/c46_ncu_launch.log 2>&1; echo "launch list rc=$?"
python tools/launch_shares.py gpurun_out/c46_launches.csv > gpurun_out/c46_launch_shares.txt 2>&1; head -12 gpurun_out/c46_launch_shares.txt

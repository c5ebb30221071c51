#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c21_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/c21_gpu_suite.log
timeout -k 5 1500 python -m pytest tests/test_full_parity_gpu.py -q -s -p no:cacheprovider -k "K2" > gpurun_out/c21_parity_k2.log 2>&1; echo "k2 rc=$?"; grep -E "flips|passed|failed" gpurun_out/c21_parity_k2.log
timeout -k 5 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c21_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c21_smoke.log
timeout -k 5 1500 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/c21_bench.err; echo "bench rc=$?"
timeout -k 5 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/c21_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu.json') if l.startswith('{')][-1])
print('1 GPU', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['frac_of_burst_peak'],3), d['clocks'], d['cpu_baseline'])
for k,v in d['extras'].items(): print(' ', k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if not isinstance(vv,(dict,list,str))})
r=json.loads([l for l in open('gpurun_out/r02_bench_reference_arm.json') if l.startswith('{')][-1])
print('ref arm', r['value'], r['cpu_baseline'])
P

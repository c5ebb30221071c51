#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c9_status.txt
for m in 0 2 1; do
B200NS_CL2=$m timeout -k 5 180 python tools/check_cg2.py 64 > gpurun_out/c9_cg2_mode$m.log 2>&1; echo "mode $m rc=$?" >> gpurun_out/c9_status.txt
done
python - <<'PY' >> gpurun_out/c9_status.txt 2>&1
import torch
a=torch.load('gpurun_out/cg2_out_mode0.pt')
for m in (1,2):
    try:
        b=torch.load(f'gpurun_out/cg2_out_mode{m}.pt')
        print('mode',m,'equal',torch.equal(a,b),'maxdiff',float((a-b).abs().max()),'rel',float((a-b).norm()/a.norm()))
    except Exception as e: print('mode',m,'ERR',e)
PY
B200NS_CL2=2 timeout -k 5 600 python bench.py --quick --no-cpu-baseline --escalate 0 > gpurun_out/c9_bench_cg2.json 2> gpurun_out/c9_bench_cg2.err; echo "bench cg2 rc=$?" >> gpurun_out/c9_status.txt
timeout -k 5 600 python bench.py --quick --no-cpu-baseline --escalate 0 > gpurun_out/c9_bench_base.json 2> gpurun_out/c9_bench_base.err; echo "bench base rc=$?" >> gpurun_out/c9_status.txt
cat gpurun_out/c9_status.txt; tail -3 gpurun_out/c9_cg2_mode2.log | cut -c1-600

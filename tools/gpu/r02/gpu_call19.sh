#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c19_det_a.log 2>&1; echo "a rc=$?"
timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c19_det_b.log 2>&1; echo "b rc=$?"
B200NS_PREC_PDL=0 timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c19_det_nopdl.log 2>&1; echo "nopdl rc=$?"
B200NS_CL2=0 timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c19_det_cl2_0.log 2>&1; echo "cl2=0 rc=$?"
for f in a b nopdl cl2_0; do tail -n 1 gpurun_out/c19_det_$f.log; done
diff gpurun_out/c19_det_a.log gpurun_out/c19_det_b.log > /dev/null && echo "a == b" || echo "a != b"
diff gpurun_out/c19_det_a.log gpurun_out/c19_det_nopdl.log > /dev/null && echo "a == nopdl" || echo "a != nopdl"
diff gpurun_out/c19_det_a.log gpurun_out/c19_det_cl2_0.log > /dev/null && echo "a == cl2_0" || echo "a != cl2_0"
timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c19_bench_quick.json 2> gpurun_out/c19_bench_quick.err; echo "bench rc=$?"
B200NS_PREC_PDL=0 timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c19_bench_quick_nopdl.json 2> gpurun_out/c19_bench_quick_nopdl.err; echo "bench nopdl rc=$?"
timeout -k 5 900 python bench.py --quick --no-cpu-baseline > gpurun_out/c19_bench_quick2.json 2> gpurun_out/c19_bench_quick2.err; echo "bench2 rc=$?"
python - <<'P'
import json
for f in ('c19_bench_quick','c19_bench_quick_nopdl','c19_bench_quick2'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],2), 'no_esc', round(d['extras']['no_escalation']['value'],1), d['escalation']['rows_refined_per_step'])
P

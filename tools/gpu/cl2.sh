#!/bin/bash
mkdir -p gpurun_out
for P in 0 1 0 1; do
B200NS_CL2=$P timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cl2_$P.json 2> gpurun_out/bench_cl2_$P.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_cl2_$P.json').read().strip().splitlines()[-1]);print('CL2=$P', round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'gemm ms',round(d['roofline']['gemm_ms_per_nfe'],3),d['clocks']['sm_mhz'])"
done
B200NS_CL2=1 timeout -k 10 600 python -m pytest tests/test_unet_gpu.py tests/test_search_gpu.py tests/test_full_size_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_cl2.log 2>&1; echo "pytest(cl2) exit $?"; tail -n 4 gpurun_out/pytest_cl2.log

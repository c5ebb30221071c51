#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
B200NS_PDL=0 timeout -k 10 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench nopdl exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_nopdl.json').read().strip().splitlines()[-1]);print('noPDL',d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
timeout -k 10 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench pdl exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl.json').read().strip().splitlines()[-1]);print('PDL',d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
B200NS_PDL=0 timeout -k 10 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl2.json 2> gpurun_out/bench_nopdl2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_nopdl2.json').read().strip().splitlines()[-1]);print('noPDL',d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
timeout -k 10 600 python bench.py --steps 18 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl2.json 2> gpurun_out/bench_pdl2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl2.json').read().strip().splitlines()[-1]);print('PDL',d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
tail -n 3 gpurun_out/bench_pdl.err

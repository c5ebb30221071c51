#!/bin/bash
# final evidence of the round (no ncu): GPU tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; head -c 300 gpurun_out/bench_ref.json; echo
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'],d['clocks'],d['cpu_baseline'])"; tail -n 3 gpurun_out/bench.err

#!/bin/bash
# Full GPU check: pytest -m gpu, smoke(), bench (b200 + reference arm), launch list under ncu.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps 18 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c18_det_a.log 2>&1; echo "a rc=$?"
timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c18_det_b.log 2>&1; echo "b rc=$?"
B200NS_PREC_PDL=0 timeout -k 5 600 python tools/check_determinism.py > gpurun_out/c18_det_nopdl.log 2>&1; echo "nopdl rc=$?"
timeout -k 5 600 python tools/check_determinism.py --escalate 0 > gpurun_out/c18_det_noesc_a.log 2>&1; echo "noesc a rc=$?"
timeout -k 5 600 python tools/check_determinism.py --escalate 0 > gpurun_out/c18_det_noesc_b.log 2>&1; echo "noesc b rc=$?"
tail -1 gpurun_out/c18_det_*.log
diff gpurun_out/c18_det_a.log gpurun_out/c18_det_b.log > /dev/null && echo "a == b" || echo "a != b"
diff gpurun_out/c18_det_a.log gpurun_out/c18_det_nopdl.log > /dev/null && echo "a == nopdl" || echo "a != nopdl"
diff gpurun_out/c18_det_noesc_a.log gpurun_out/c18_det_noesc_b.log > /dev/null && echo "noesc a == b" || echo "noesc a != b"

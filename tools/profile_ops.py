"""Per-op CUDA-event timing of one ADM-64 NFE at batch B (default 64): TFLOP/s for GEMM/attention ops,
GB/s for the HBM-bound ops.  Usage: python tools/profile_ops.py [B] [--csv path]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.unet import UNetEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
eng = UNetEngine(random_state_dict(adm_param_shapes(), 1234), device='cuda', use_graphs='--graph' in sys.argv,
                 alternate_walk='--no-alt' not in sys.argv, fused_gn_stats='--no-fused-stats' not in sys.argv,
                 lanes=int(sys.argv[sys.argv.index('--lanes') + 1]) if '--lanes' in sys.argv else None,
                 lane_min_res=int(sys.argv[sys.argv.index('--lane-min-res') + 1]) if '--lane-min-res' in sys.argv else 32)
fp = eng.plan(B, 1)
fp.x_in.normal_()
plan = fp.plan
for _ in range(2):
    plan.run_timed()
runs = [plan.run_timed() for _ in range(5)]
ms = [min(r[i] for r in runs) for i in range(len(runs[0]))]
rows = []
for i, (lab, kind, fl, t) in enumerate(zip(plan.labels, plan.kinds, plan.flops, ms)):
    rows.append((i, lab, kind, t, fl / (t * 1e-3) / 1e12 if fl else 0.0))
tot = sum(ms)
print(f'B={B} ops={len(ms)} total {tot:.3f} ms')
if '--graph' in sys.argv:
    for _ in range(3): plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.run()
    e1.record(); torch.cuda.synchronize()
    print(f'graph NFE: {e0.elapsed_time(e1) / 10:.3f} ms')
agg = {}
for i, lab, kind, t, tf in rows:
    a = agg.setdefault(kind, [0.0, 0.0, 0])
    a[0] += t; a[1] += plan.flops[i]; a[2] += 1
for k, (t, fl, n) in agg.items():
    print(f'  {k:10s} n={n:4d} {t:8.3f} ms  {fl / (t * 1e-3) / 1e12 if fl else 0:8.1f} TFLOP/s')
print('--- slowest 40 ops')
for i, lab, kind, t, tf in sorted(rows, key=lambda r: -r[3])[:40]:
    print(f'{i:4d} {kind:9s} {t * 1e3:9.1f} us  {tf:7.1f} TF  {lab}')
if '--csv' in sys.argv:
    with open(sys.argv[sys.argv.index('--csv') + 1], 'w') as f:
        f.write('idx,label,kind,us,tflops\n')
        for i, lab, kind, t, tf in rows:
            f.write(f'{i},{lab},{kind},{t * 1e3:.2f},{tf:.1f}\n')

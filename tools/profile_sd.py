"""SD-1.5-shaped UNet engine (config 5): per-op CUDA-event timing of one forward at UNet batch M (default 64 = 32
candidates x 2 CFG halves) and the time of whole beam steps.
Usage: python tools/profile_sd.py [M] [--beam B N STEPS] [--csv path]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200 import ops
from diffusion_tts_b200.arch import sd_unet_param_shapes, random_state_dict
from diffusion_tts_b200.sd_unet import SDUNetEngine
from diffusion_tts_b200.sd.beam import DDIMTable, sd_beam_search

M = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
t0 = time.time()
eng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device='cuda', use_graphs='--graph' in sys.argv)
eng.set_context(torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(1)).cuda())
print(f'engine built in {time.time() - t0:.1f} s')
fp = eng.plan(M, 64)
fp.x_in.normal_()
fp.emb_in.copy_(eng.timestep_embedding(500))
plan = fp.plan
for _ in range(2):
    plan.run_timed()
runs = [plan.run_timed() for _ in range(3)]
ms = [min(r[i] for r in runs) for i in range(len(runs[0]))]
tot = sum(ms)
fl = sum(plan.flops)
print(f'M={M} ops={len(ms)} total {tot:.3f} ms  GEMM flops {fl / 1e12:.2f} TFLOP -> {fl / (tot * 1e-3) / 1e12:.1f} TFLOP/s whole forward; '
      f'mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB')
agg = {}
for i, (kind, t) in enumerate(zip(plan.kinds, ms)):
    a = agg.setdefault(kind, [0.0, 0.0, 0])
    a[0] += t; a[1] += plan.flops[i]; a[2] += 1
for k, (t, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:12s} n={n:4d} {t:8.3f} ms {100 * t / tot:5.1f}%  {f / (t * 1e-3) / 1e12 if f else 0:8.1f} TFLOP/s')
print('--- slowest 30 ops')
rows = [(i, plan.labels[i], plan.kinds[i], ms[i], plan.flops[i] / (ms[i] * 1e-3) / 1e12 if plan.flops[i] else 0.0) for i in range(len(ms))]
for i, lab, kind, t, tf in sorted(rows, key=lambda r: -r[3])[:30]:
    print(f'{i:4d} {kind:10s} {t * 1e3:9.1f} us  {tf:7.1f} TF  {lab}')
if '--csv' in sys.argv:
    with open(sys.argv[sys.argv.index('--csv') + 1], 'w') as f:
        f.write('idx,label,kind,us,tflops\n')
        for i, lab, kind, t, tf in rows:
            f.write(f'{i},{lab},{kind},{t * 1e3:.2f},{tf:.1f}\n')
if '--beam' in sys.argv:
    j = sys.argv.index('--beam')
    B, N, S = int(sys.argv[j + 1]), int(sys.argv[j + 2]), int(sys.argv[j + 3])
    tab = DDIMTable(50)
    lat = torch.randn(1, 4, 64, 64).cuda()
    sd_beam_search(eng, tab, lat, None, B, N, steps=[0])                  # warm-up (plans, graphs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count() if hasattr(ops, 'launch_count') else 0
    e0.record()
    _, rec = sd_beam_search(eng, tab, lat, None, B, N, steps=list(range(S)))
    e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / S
    print(f'beam B={B} N={N}: {dt:.1f} ms/step, {B * N / dt * 1e3:.1f} scored candidates/s, '
          f'{(2 * B * N + 2 * B) / dt * 1e3:.1f} UNet forwards/s; mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB')

"""SD-1.5 VAE decoder: per-op CUDA-event timing of one decode of B latents 64x64 -> 512x512, and the time of a
beam step with decode-then-score.  Usage: python tools/profile_vae.py [B] [--beam B N STEPS] [--csv path]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes, vae_decoder_param_shapes
from diffusion_tts_b200.vae import DecodedImageScorer, VAEDecoderEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
eng = VAEDecoderEngine(random_state_dict(vae_decoder_param_shapes(), 4321), device='cuda', use_graphs='--graph' in sys.argv)
fp = eng.plan(B, 64)
fp.z.normal_()
plan = fp.plan
for _ in range(2):
    plan.run_timed()
runs = [plan.run_timed() for _ in range(3)]
ms = [min(r[i] for r in runs) for i in range(len(runs[0]))]
tot, fl = sum(ms), sum(plan.flops)
print(f'B={B} ops={len(ms)} total {tot:.3f} ms ({tot / B:.3f} ms per image)  GEMM flops {fl / 1e12:.2f} TFLOP -> '
      f'{fl / (tot * 1e-3) / 1e12:.1f} TFLOP/s whole decode; mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB')
agg = {}
for i, (kind, t) in enumerate(zip(plan.kinds, ms)):
    a = agg.setdefault(kind, [0.0, 0.0, 0])
    a[0] += t; a[1] += plan.flops[i]; a[2] += 1
for k, (t, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:12s} n={n:4d} {t:8.3f} ms {100 * t / tot:5.1f}%  {f / (t * 1e-3) / 1e12 if f else 0:8.1f} TFLOP/s')
rows = [(i, plan.labels[i], plan.kinds[i], ms[i], plan.flops[i] / (ms[i] * 1e-3) / 1e12 if plan.flops[i] else 0.0) for i in range(len(ms))]
print('--- slowest 25 ops')
for i, lab, kind, t, tf in sorted(rows, key=lambda r: -r[3])[:25]:
    print(f'{i:4d} {kind:12s} {t * 1e3:9.1f} us  {tf:7.1f} TF  {lab}')
if '--csv' in sys.argv:
    with open(sys.argv[sys.argv.index('--csv') + 1], 'w') as f:
        f.write('idx,label,kind,us,tflops\n')
        for i, lab, kind, t, tf in rows:
            f.write(f'{i},{lab},{kind},{t * 1e3:.2f},{tf:.1f}\n')
if '--beam' in sys.argv:
    from diffusion_tts_b200.sd.beam import DDIMTable, sd_beam_search
    from diffusion_tts_b200.sd_unet import SDUNetEngine
    j = sys.argv.index('--beam')
    Bm, N, S = int(sys.argv[j + 1]), int(sys.argv[j + 2]), int(sys.argv[j + 3])
    ueng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device='cuda')
    ueng.set_context(torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(1)).cuda())
    tab = DDIMTable(50)
    lat = torch.randn(1, 4, 64, 64).cuda()
    scorer = DecodedImageScorer(eng, None, chunk=B)
    kw = dict(decode=lambda x: x, scorer=scorer)
    sd_beam_search(ueng, tab, lat, None, Bm, N, steps=[0], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sd_beam_search(ueng, tab, lat, None, Bm, N, steps=list(range(S)), **kw)
    e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / S
    print(f'beam B={Bm} N={N} with VAE decode-then-score: {dt:.1f} ms/step, {Bm * N / dt * 1e3:.1f} scored candidates/s; '
          f'mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB')

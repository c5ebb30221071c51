"""CLIP ViT-L/14 scorer: per-op CUDA-event timing of scoring B 512x512 uint8 images, and the graph time.
Usage: python tools/profile_clip.py [B] [--csv path]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import clip_param_shapes, random_state_dict
from diffusion_tts_b200.clip import CLIPVisionEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
sd = random_state_dict(clip_param_shapes(vision_only=True), 99)
eng = CLIPVisionEngine(sd, device='cuda', use_graphs=False)
cp = eng.plan(B, 512, 512)
cp.images.random_(0, 256)
cp.text.normal_()
plan = cp.plan
for _ in range(2):
    plan.run_timed()
runs = [plan.run_timed() for _ in range(3)]
ms = [min(r[i] for r in runs) for i in range(len(runs[0]))]
tot, fl = sum(ms), sum(plan.flops)
real = 2.0 * B * eng.tokens * sum(sd[k].numel() for k in sd if k.endswith('proj.weight') or 'mlp.fc' in k and k.endswith('weight'))
print(f'B={B} ops={len(ms)} total {tot:.3f} ms ({tot / B:.3f} ms per image)  executed GEMM+attention flops {fl / 1e12:.2f} TFLOP -> '
      f'{fl / (tot * 1e-3) / 1e12:.1f} TFLOP/s (rows padded {eng.tokens} -> {eng.Lp}); finite {bool(torch.isfinite(cp.scores).all())}')
agg = {}
for i, (kind, t) in enumerate(zip(plan.kinds, ms)):
    a = agg.setdefault(kind, [0.0, 0.0, 0])
    a[0] += t; a[1] += plan.flops[i]; a[2] += 1
for k, (t, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:16s} n={n:4d} {t:8.3f} ms {100 * t / tot:5.1f}%  {f / (t * 1e-3) / 1e12 if f else 0:8.1f} TFLOP/s')
rows = [(i, plan.labels[i], plan.kinds[i], ms[i], plan.flops[i] / (ms[i] * 1e-3) / 1e12 if plan.flops[i] else 0.0) for i in range(len(ms))]
print('--- layer 0 + head ops')
for i, lab, kind, t, tf in rows[:10] + rows[-3:]:
    print(f'{i:4d} {kind:16s} {t * 1e3:9.1f} us  {tf:7.1f} TF  {lab}')
torch.cuda.synchronize()
plan.instantiate_graph()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    plan.run()
e0.record()
for _ in range(10):
    plan.run()
e1.record(); torch.cuda.synchronize()
g = e0.elapsed_time(e1) / 10
print(f'graph: {g:.3f} ms per batch of {B} = {B / g * 1e3:.0f} images/s')
if '--csv' in sys.argv:
    with open(sys.argv[sys.argv.index('--csv') + 1], 'w') as f:
        f.write('idx,label,kind,us,tflops\n')
        for i, lab, kind, t, tf in rows:
            f.write(f'{i},{lab},{kind},{t * 1e3:.2f},{tf:.1f}\n')

"""Run selected ops of one ADM-64 NFE plan (batch B; `--sd`: the SD-1.5-shaped UNet forward at UNet batch B) in isolation,
inside a cudaProfilerStart/Stop window:
  ncu --profile-from-start off --set full ... python tools/profile_one.py 64 enc.32x32_block0.qkv dec.32x32_block0.proj
  ncu --profile-from-start off --set full ... python tools/profile_one.py 64 --sd up_blocks.3.attentions.1.transformer_blocks.0.attn1.attn
Without ncu it prints the CUDA-event time of each selected op (best of 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.unet import UNetEngine

B = int(sys.argv[1])
labels = [a for a in sys.argv[2:] if a != '--sd']
if '--sd' in sys.argv:
    from diffusion_tts_b200.arch import sd_unet_param_shapes
    from diffusion_tts_b200.sd_unet import SDUNetEngine
    eng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device='cuda', use_graphs=False)
    eng.set_context(torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(1)).cuda())
    fp = eng.plan(B, 64)
    fp.emb_in.copy_(eng.timestep_embedding(500))
else:
    eng = UNetEngine(random_state_dict(adm_param_shapes(), 1234), device='cuda', use_graphs=False)
    fp = eng.plan(B, 1)
fp.x_in.normal_()
plan = fp.plan
plan.run()
torch.cuda.synchronize()
idx = [next(i for i, x in enumerate(plan.labels) if x == l or x.startswith(l + '[')) for l in labels]      # 'name[cols]': slices / phases
for i in idx:                                   # warm
    plan.run(i, i + 1)
torch.cuda.synchronize()
for i in idx:
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run(i, i + 1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = plan.flops[i]
    print(f'{plan.labels[i]:36s} {best * 1e3:9.1f} us  {fl / (best * 1e-3) / 1e12 if fl else 0:7.1f} TFLOP/s')
torch.cuda.profiler.start()
for i in idx:
    plan.run(i, i + 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

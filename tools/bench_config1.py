"""BASELINE.json configs[0]: EDM CIFAR-10 32x32 DDPM++ (SongUNet, 55.7 M parameters, random init), --method naive, 18-step
stochastic Heun, brightness scorer, batch 1 (the reference's own CPU-runnable case) -- latency of one image on a B200, and the
throughput of the same sampler at batch 64.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import ddpmpp_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
from diffusion_tts_b200.edm.main import naive_search
from diffusion_tts_b200.scorers import BrightnessScorer
from diffusion_tts_b200 import ops

net = B200Denoiser(random_state_dict(ddpmpp_param_shapes(), 4321), device='cuda')
table = StepTable(net, 'cuda', 18, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
scorer = BrightnessScorer(device='cuda')
out = {}
for B in (1, 64):
    lat = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(3):
        naive_search(net, lat, None, table)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        x, _ = naive_search(net, lat, None, table)
        s = scorer(ops.quantize_u8(x.contiguous()), None, None)
    e1.record()
    torch.cuda.synchronize()
    out[B] = e0.elapsed_time(e1) / reps
print(json.dumps({'metric': 'ms_per_image (config 1: DDPM++ 32x32 naive, 18 Heun steps = 35 NFE, brightness)', 'batch1_ms': out[1],
                  'batch64_ms': out[64], 'batch64_images_per_sec': 64 / (out[64] / 1e3), 'nfe_per_image': 35,
                  'gflop_per_image': 35 * 42.38}))

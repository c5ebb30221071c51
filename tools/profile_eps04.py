"""Per-step device time of the eps = 0.4 extra of bench.py (same seeds and noise), with and without escalation, to see
where its ms per step go.  Usage: PYTHONHASHSEED=0 python tools/profile_eps04.py"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
from diffusion_tts_b200.edm.main import SamplingParams, eps_greedy_search
from diffusion_tts_b200.scorers import BrightnessScorer
import bench as B

dev = torch.device('cuda', 0)
N = 64
net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
table = StepTable(net, dev, B.NUM_STEPS, **B.SAMPLER)
scorer = BrightnessScorer(device=dev)
g = torch.Generator().manual_seed(1)
latents = torch.randn(1, 3, 64, 64, generator=g)
labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
order = list(range(B.NUM_STEPS))
noise = {}
for i in order:
    noise[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
    noise[i] = torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
x0 = (latents.to(torch.float64) * table.t_steps[0].cpu()).to(dev)
gd = torch.Generator(device=dev).manual_seed(5)
noise04 = dict(noise)
for i in order:
    t = torch.randn(N, 1, 3, 64, 64, generator=gd, dtype=torch.float64, device=dev)
    for n in range(N):
        noise04[f'fresh_{i}_0_{n}'] = t[n]


def run(eps, nz, escalate):
    p = SamplingParams(N=N, K=1, eps=eps, lambda_param=0.15, scorer=scorer)
    out = None
    for rep in range(2):                       # second pass is the measured one
        torch.manual_seed(11)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(order) + 1)]
        host = [time.perf_counter()]
        torch.cuda.synchronize()
        evs[0].record()

        def on_step(i, x_next, idx, scores):
            evs[len(host)].record()
            host.append(time.perf_counter())
        x, rec = eps_greedy_search(net, None, labels, p, table, precomputed_noise=nz, step_indices=order, x_init=x0,
                                   on_step=on_step, escalate=escalate)
        torch.cuda.synchronize()
        out = dict(eps=eps, escalate=escalate, total_ms=round(evs[0].elapsed_time(evs[-1]), 2),
                   per_step_ms=[round(evs[j].elapsed_time(evs[j + 1]), 2) for j in range(len(order))],
                   host_ms=[round((host[j + 1] - host[j]) * 1e3, 2) for j in range(len(order))], rows=rec.escalated)
    print(json.dumps(out), flush=True)


for eps, nz in ((0.0, noise), (0.4, noise04)):
    for esc in (False, True):
        run(eps, nz, esc)

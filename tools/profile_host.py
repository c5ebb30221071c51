"""Host-side cost of the search loop: cProfile of eps_greedy steps (no syncs inside)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
from diffusion_tts_b200.edm.main import SamplingParams, eps_greedy_search
from diffusion_tts_b200.scorers import BrightnessScorer
dev = torch.device('cuda')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
table = StepTable(net, dev, 18, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
g = torch.Generator().manual_seed(1)
labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
x = (torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) * 80).to(dev)
steps = [6, 7, 8, 9]
pre = {}
for i in steps:
    pre[i] = torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
    pre[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
params = SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=BrightnessScorer(device=dev))
run = lambda: eps_greedy_search(net, None, labels, params, table, precomputed_noise=pre, step_indices=steps, x_init=x)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'N={N}: host {1e3 * (t1 - t0) / len(steps):.2f} ms/step, host+gpu {1e3 * (t2 - t0) / len(steps):.2f} ms/step')
pr = cProfile.Profile(); pr.enable(); run(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

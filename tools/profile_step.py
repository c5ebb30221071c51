"""Where does a search step go?  CUDA-event timing of the phases of one eps_greedy step (N=64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200 import ops
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable, HeunStepper
from diffusion_tts_b200.edm.main import SamplingParams, eps_greedy_search
from diffusion_tts_b200.scorers import BrightnessScorer

dev = torch.device('cuda')
N = 64
net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
table = StepTable(net, dev, 18, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
g = torch.Generator().manual_seed(1)
labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
x = (torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) * 80).to(dev)
cands = torch.randn(N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
stepper = HeunStepper(net, table, labels)
fp = net.engine.plan(N, 1)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print('NFE graph launch (B=64)      : %.3f ms' % timeit(lambda: fp.plan.run()))
print('full candidate step (2 NFE)  : %.3f ms' % timeit(lambda: stepper.step(x, cands, 8, want_x_next=True, want_sums=True)))
c = table.steps[8]
print('heun_pre                     : %.3f ms' % timeit(lambda: ops.heun_pre(x, cands, c.s, c.c_in1, net_in=fp.x_in)))
xh, _ = ops.heun_pre(x, cands, c.s, c.c_in1, net_in=fp.x_in)
F1 = fp.out.clone()
print('F1 clone                     : %.3f ms' % timeit(lambda: fp.out.clone()))
print('heun_mid                     : %.3f ms' % timeit(lambda: ops.heun_mid(xh, F1, c.c_skip1, c.c_out1, c.t_hat, c.dt, c.c_in2, net_in2=fp.x_in)))
print('heun_post (+sums, +x_next)   : %.3f ms' % timeit(lambda: ops.heun_post(xh, F1, F1, c.c_skip1, c.c_out1, c.t_hat, c.dt, c.c_skip2, c.c_out2, c.t_next, want_x_next=True, want_sums=True)))
params = SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=BrightnessScorer(device=dev))
pre = {8: torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev), 'pivot_8': torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)}
print('eps_greedy step i=8 (total)  : %.3f ms' % timeit(lambda: eps_greedy_search(net, None, labels, params, table, precomputed_noise=pre, step_indices=[8], x_init=x), n=5))
import time
t0 = time.perf_counter()
for _ in range(5): eps_greedy_search(net, None, labels, params, table, precomputed_noise=pre, step_indices=[8], x_init=x)
print('  host time per step (no sync): %.3f ms' % ((time.perf_counter() - t0) / 5 * 1e3)); torch.cuda.synchronize()

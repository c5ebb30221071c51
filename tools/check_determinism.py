"""Run-to-run determinism of the headline search (ADM-64, N=64, 18 steps, escalation on): the same process runs it twice and
prints a digest of every round's scores, refined scores, selected indices and committed states.  Run the script twice (and
with B200NS_PREC_PDL=0 / B200NS_CL2=0) and diff the output: every digest must agree.
Usage: python tools/check_determinism.py [--escalate 0|1]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get('PYTHONHASHSEED') is None:     # the reference's candidate scales come from Python's salted hash("i_k_n")
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
from diffusion_tts_b200.edm.main import SamplingParams, eps_greedy_search
from diffusion_tts_b200.scorers import BrightnessScorer

esc = '--escalate' not in sys.argv or sys.argv[sys.argv.index('--escalate') + 1] != '0'
dev = torch.device('cuda')
net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
table = StepTable(net, dev, 18, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
N = 64
params = SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=BrightnessScorer(device=dev))
g = torch.Generator().manual_seed(1)
latents = torch.randn(1, 3, 64, 64, generator=g)
labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
noise = {}
for i in range(18):
    noise[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
    noise[i] = torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)


def digest(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:10]


def run():
    x, rec = eps_greedy_search(net, latents.to(dev), labels, params, table, precomputed_noise=noise, record=True, escalate=esc)
    torch.cuda.synchronize()
    rows = []
    for r in range(len(rec.scores)):
        ref = rec.refined[r] if r < len(rec.refined) and rec.refined[r] is not None else None
        rows.append((r, digest(rec.scores[r]), digest(ref) if ref is not None else '-', int(rec.indices[r][0]), rec.escalated[r],
                     digest(rec.x_steps[r]) if r < len(rec.x_steps) else '-'))
    return rows, digest(x)


a, xa = run()
b, xb = run()
for ra, rb in zip(a, b):
    flag = '' if ra == rb else '   <-- differs between the two in-process runs'
    print('round %2d scores %s refined %s idx %2d escalated %d x %s%s' % (*ra, flag))
print('final x', xa, xb, 'IN-PROCESS DETERMINISTIC' if (a == b and xa == xb) else 'NONDETERMINISTIC')

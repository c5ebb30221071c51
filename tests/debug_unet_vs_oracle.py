"""GPU debug: per-block comparison of the engine against the CPU oracle (tiny ADM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch
from oracle import edm_oracle as O
from tests.helpers import load_golden, oracle_net
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.denoiser import B200Denoiser

name = sys.argv[1] if len(sys.argv) > 1 else 'unet_tiny_adm.pt'
g = load_golden(name)
onet, spec, sd = oracle_net(g['cfg'], g['seed'])
case = g['cases'][1 if len(g['cases']) > 1 else 0]
x, labels = case['x'], case['labels']
sigma = torch.tensor(case['sigma'], dtype=torch.float64)
c_skip, c_out, c_in, c_noise = O.precond_coeffs(sigma)
xin = c_in * x
# oracle intermediates
emb = O.embedding(spec, sd, c_noise.flatten(), labels)
ref = {}
skips = []
h = xin
for b in spec.enc:
    h = O._conv(sd, b.name, h, 3) if b.kind == 'conv' else O.unet_block(spec, b, sd, h, emb)
    ref[b.name] = h
    skips.append(h)
for b in spec.dec:
    if b.kind != 'block':
        continue
    if h.shape[1] != b.cin:
        h = torch.cat([h, skips.pop()], dim=1)
    h = O.unet_block(spec, b, sd, h, emb)
    ref[b.name] = h
eng = B200Denoiser(sd, device='cuda')
B = x.shape[0]
D = eng(x.cuda(), sigma.cuda(), labels.cuda() if labels is not None else None)
fp = eng.engine.plan(B, eng._distinct_rows(labels.cuda(), B) if labels is not None else 1)
print('emb rel', ((fp.emb.cpu() - emb[:fp.b_emb]).norm() / emb.norm()).item())
for k, v in fp.block_out.items():
    got = v.float().cpu().permute(0, 3, 1, 2)
    r = ref[k]
    print(f'{k:28s} rel {((got - r).norm() / r.norm()).item():.4e}  max|ref| {r.abs().max().item():.3f}')
Fx = (D.cpu() - c_skip * x) / c_out
print('F rel', ((Fx - case['F']).norm() / case['F'].norm()).item())

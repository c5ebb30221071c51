"""Which layer (if any) gives different bits for the same samples evaluated at two batch sizes?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.unet import UNetEngine
B1, B2 = int(sys.argv[1]), int(sys.argv[2])
eng = UNetEngine(random_state_dict(adm_param_shapes(), 1234), device='cuda', use_graphs=False)
g = torch.Generator().manual_seed(3)
x = torch.randn(B2, 3, 64, 64, generator=g).cuda()
outs = {}
for B in (B1, B2):
    fp = eng.plan(B, 1)
    fp.x_in.copy_(x[:B])
    fp.emb_in.normal_(generator=None) if False else fp.emb_in.fill_(0.1)
    fp.labels.zero_(); fp.labels[:, 3] = 1
    fp.plan.run()
    torch.cuda.synchronize()
    outs[B] = ({k: v.clone() for k, v in fp.block_out.items()}, fp.out.clone(), fp)
n = min(B1, B2)
bad = 0
for k in outs[B1][0]:
    a, b = outs[B1][0][k][:n], outs[B2][0][k][:n]
    if not torch.equal(a, b):
        d = (a.float() - b.float()).abs().max().item()
        print('DIFF', k, tuple(a.shape), 'max abs', d)
        bad += 1
        if bad > 3: break
print('final out equal:', torch.equal(outs[B1][1][:n], outs[B2][1][:n]), 'layers differing:', bad)

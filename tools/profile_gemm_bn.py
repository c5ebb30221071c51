"""Measures the GEMM tile widths (BN 64/128/192/256) against each other on conv / 1x1 shapes whose N all widths divide.
Usage: python tools/profile_gemm_bn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusion_tts_b200._lib import ACT_DTYPE as ACT  # noqa: E402  (the engine's 16-bit storage type)
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200 import ops, _lib as L

def run(B, H, Cin, N, taps, bn, res=False, stats=False):
    L.lib().b200ns_debug_force_tile_width(bn)
    x = torch.randn(B, H, H, Cin, device='cuda').to(ACT)
    w = (torch.randn(N, taps * Cin, device='cuda') / (taps * Cin) ** 0.5).to(ACT)
    out = torch.empty(B, H, H, N, device='cuda', dtype=ACT)
    r = torch.randn(B, H, H, N, device='cuda').to(ACT) if res else None
    st = torch.empty(B * H * H // 64, N, 2, device='cuda') if stats else None
    plan = ops.Plan()
    plan.add_gemm([x], [(0, taps, 0, Cin // 64)], w, N, out, residual=r, gn_stats=st)
    L.lib().b200ns_debug_force_tile_width(0)
    for _ in range(3):
        plan.run()
    ts = []
    for _ in range(7):
        ts.append(sum(plan.run_timed()))
    t = sorted(ts)[len(ts) // 2]
    return t, 2.0 * B * H * H * N * taps * Cin / (t * 1e-3) / 1e12

shapes = [(64, 64, 384, 768, 9), (64, 32, 768, 768, 9), (64, 16, 1536, 1536, 9), (64, 8, 1536, 1536, 9),
          (64, 64, 384, 3072, 1), (64, 32, 768, 768, 1), (64, 32, 1536, 768, 9)]
for (B, H, Cin, N, taps) in shapes:
    line = f'B={B} H={H} Cin={Cin} N={N} taps={taps}: '
    for bn in (64, 128, 192, 256):
        t, tf = run(B, H, Cin, N, taps, bn, res=True, stats=True)
        line += f' BN{bn} {t * 1e3:7.1f}us {tf:6.0f}TF |'
    print(line, flush=True)

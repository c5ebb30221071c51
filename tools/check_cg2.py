"""A/B of the GEMM launch modes (B200NS_CL2 = 0 | 1 | 2): one ADM-64 forward at batch 64 on fixed inputs; saves the output
for a bit-level comparison across modes and prints per-kind timings.  Builder tool (gpurun)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge
    ge.build(oracle=False)
    from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
    from diffusion_tts_b200.denoiser import B200Denoiser
    mode = os.environ.get('B200NS_CL2', '0')
    net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device='cuda')
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    fp = net.engine.plan(B, 1)
    g = torch.Generator(device='cuda').manual_seed(5)
    fp.x_in.copy_(torch.randn(fp.x_in.shape, device='cuda', generator=g))
    fp.emb_in.copy_(torch.randn(fp.emb_in.shape, device='cuda', generator=g))
    fp.labels.zero_()
    fp.labels[:, 3] = 1
    for _ in range(3):
        fp.plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fp.plan.run()
    e1.record()
    torch.cuda.synchronize()
    per_op = fp.plan.run_timed()
    by = {}
    for t, k in zip(per_op, fp.plan.kinds):
        by[k] = by.get(k, 0.0) + t
    out = fp.out.clone()
    torch.save(out.cpu(), os.path.join(ROOT, 'gpurun_out', f'cg2_out_mode{mode}.pt'))
    gemm_flops = sum(f for f, k in zip(fp.plan.flops, fp.plan.kinds) if k == 'gemm')
    print(json.dumps({'mode': mode, 'B': B, 'nfe_ms_graph': e0.elapsed_time(e1) / 10, 'by_kind': by, 'finite': bool(torch.isfinite(out).all()),
                      'gemm_tflops': gemm_flops / by['gemm'] / 1e9, 'checksum': float(out.double().abs().sum())}))
    top = sorted(((t, l) for t, l, k in zip(per_op, fp.plan.labels, fp.plan.kinds) if k == 'gemm'), reverse=True)[:6]
    print('   slowest gemms:', [(round(t, 3), l) for t, l in top])


if __name__ == '__main__':
    main()

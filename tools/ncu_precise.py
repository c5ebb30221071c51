"""One eager pass of the precise (split-fp16) plan at batch R, for `ncu --metrics gpu__time_duration.sum`: the per-kernel
durations show where a contender re-evaluation spends its time.  Builder tool (gpurun)."""
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
    from diffusion_tts_b200.precise import PreciseUNetEngine
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sd = random_state_dict(adm_param_shapes(), 1234)
    net = B200Denoiser(sd, device='cuda')
    eng = PreciseUNetEngine(net.engine, sd, use_graphs=False)
    fp = eng.plan(R, 1)
    fp.x_in.normal_()
    fp.plan.run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push('precise_pass')
    fp.plan.run()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print('labels', len(fp.plan.labels))
    with open(os.path.join(ROOT, 'gpurun_out', f'precise_labels_R{R}.txt'), 'w') as f:
        f.write('\n'.join(f'{k}\t{l}' for k, l in zip(fp.plan.kinds, fp.plan.labels)))


if __name__ == '__main__':
    main()

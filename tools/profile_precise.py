"""Time the fp32-faithful (split-fp16) re-scoring engine per contender batch size: CUDA-event time of one precise NFE (graph
launch) and the per-op split, next to one bf16 NFE at batch 64.  Prints JSON lines; builder tool (gpurun)."""
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
    net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device='cuda')
    eng = net.precise_engine
    for R in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
        fp = eng.plan(R, 1)
        fp.x_in.normal_()
        for _ in range(3):
            fp.plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fp.plan.run()
        e1.record()
        torch.cuda.synchronize()
        per_op = fp.plan.run_timed()
        by = {}
        for t, k in zip(per_op, fp.plan.kinds):
            by[k] = by.get(k, 0.0) + t
        flops = sum(fp.plan.flops)
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({'precise_R': R, 'nfe_ms_graph': ms, 'nfe_ms_eager_sum': sum(per_op), 'by_kind': by,
                          'mma_tflops': flops / ms / 1e9, 'finite': bool(torch.isfinite(fp.out).all())}))
        top = sorted(zip(per_op, fp.plan.labels), reverse=True)[:8]
        print('   slowest ops:', [(round(t, 3), l) for t, l in top])


if __name__ == '__main__':
    main()

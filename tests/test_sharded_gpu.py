"""Candidate sharding on the GPU: sharded == unsharded, BIT-identical.

Two processes share cuda:0 (the round-end GPU box has one device, and NCCL refuses two ranks on one device), joined by a
`gloo` process group that all-reduces the CUDA tensors of the search loop: each rank evaluates its half of the candidates
with the sm_100a kernels, the packed (score, index) key is max-reduced, the winner is exchanged.  Rank 0 also runs the same
search unsharded and compares: selected indices, committed noise and committed trajectory must be identical bit for bit --
with near-tie escalation on and off (the contenders of one image may sit on different ranks).
(8-GPU NCCL runs of the same check: tools/check_sharded.py, profiles/r0*_sharded_*.json.)
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import edm_oracle as O
    from diffusion_tts_b200 import build
    build.build()
    import diffusion_tts_b200.denoiser as den
    import diffusion_tts_b200.edm.main as em
    import diffusion_tts_b200.scorers as sc
    try:
        torch.cuda.set_device(0)
        dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
        cfg = dict(model_type='DhariwalUNet', img_resolution=16, in_channels=3, out_channels=3, label_dim=10,
                   model_channels=64, channel_mult=[1, 2], num_blocks=1, attn_resolutions=[8])
        spec = O.build_unet_spec(**cfg)
        sd = O.seeded_state_dict(O.unet_param_shapes(spec), 11)
        net = den.B200Denoiser(sd, device='cuda')
        steps, N, K, b = 6, 8, 2, 2
        kw = dict(S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
        table = den.StepTable(net, 'cuda', steps, **kw)
        g = torch.Generator().manual_seed(123)
        latents = torch.randn(b, 3, 16, 16, generator=g).cuda()
        labels = torch.eye(10)[torch.randint(10, (b,), generator=g)].cuda()
        pre = {}
        for i in range(steps):
            pre[f'pivot_{i}'] = torch.randn(b, 3, 16, 16, generator=g, dtype=torch.float64).cuda()
            pre[i] = torch.randn(b, K, N, 3, 16, 16, generator=g, dtype=torch.float64).cuda()
        scales = (torch.arange(steps * K * N, dtype=torch.float32).reshape(steps, K, N) * 0.37 % 1.0) * 16.0
        params = em.SamplingParams(N=N, K=K, eps=0.0, lambda_param=0.15, scorer=sc.BrightnessScorer(device='cuda'))
        out = {}
        for esc, delta in ((False, 0.0), (True, 5e-3)):          # a wide delta: several contenders in most rounds
            run = lambda sh: em.eps_greedy_search(net, latents, labels, params, table, precomputed_noise=pre, shard=sh,
                                                  record=True, scale_table=scales, escalate=esc, delta=delta)
            x_sh, rec_sh = run(em.Shard(rank, world, None))
            torch.cuda.synchronize()
            ok, n_esc = True, sum(rec_sh.escalated)
            if rank == 0:
                x_1, rec_1 = run(em.Shard())
                torch.cuda.synchronize()
                ok = (torch.equal(torch.stack(rec_sh.indices), torch.stack(rec_1.indices)) and torch.equal(x_sh, x_1) and
                      all(torch.equal(a, c) for a, c in zip(rec_sh.x_steps, rec_1.x_steps)) and
                      all(torch.equal(a, c) for a, c in zip(rec_sh.pivots, rec_1.pivots)))
                n_esc = (n_esc, sum(rec_1.escalated))
            out[esc] = (ok, n_esc)
            dist.barrier()
        q.put((rank, out, None))
        dist.destroy_process_group()
    except Exception as e:          # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, None, traceback.format_exc()))


def test_sharded_search_equals_unsharded_bit_for_bit():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, out, err in res:
        assert err is None, err
        for esc, (ok, n_esc) in out.items():
            assert ok, (rank, esc, n_esc)
            if esc and rank == 0:
                assert n_esc[1] > 0, 'the escalated run must actually have refined some contenders'

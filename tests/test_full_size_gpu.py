"""Full-size property tests at BASELINE.json's shapes (ImageNet-64 ADM, 295.9 M parameters, N = 64 candidates), where the
fp32 CPU oracle would take minutes per step: size-independent properties of the candidate-batched search step.

  * determinism: the same inputs give the same bits twice (no floating-point atomics anywhere on the path);
  * candidate independence: a candidate's score does not depend on its position in the batch or on its neighbours --
    permuting the directions permutes the scores bit-exactly, and argmax follows the permutation (first-max rule);
  * duplicated candidates tie exactly and the lower index wins (edm/main.py:842, torch.argmax);
  * noise-free steps (gamma = 0): all N candidates identical => N-way exact tie => index 0;
  * candidate geometry: ||cand - pivot||_2 == h * lambda * sqrt(3*64*64) (edm/main.py:764-779) to fp64 rounding;
  * the committed state equals the winning candidate's own Heun result (commit reuse == recompute, bit-identical).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, STEPS = 64, 18
SAMPLER = dict(S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)


@pytest.fixture(scope='module')
def ctx():
    from diffusion_tts_b200 import build
    build.build()
    from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
    from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
    import diffusion_tts_b200.edm.main as em
    from diffusion_tts_b200.scorers import BrightnessScorer
    net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device='cuda')
    table = StepTable(net, 'cuda', STEPS, **SAMPLER)
    g = torch.Generator().manual_seed(3)
    labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].cuda()
    x0 = (torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) * 5).cuda()
    params = em.SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=BrightnessScorer(device='cuda'))
    return dict(net=net, table=table, em=em, labels=labels, x0=x0, params=params, g=g)


def _noise(g, steps, dirs=None):
    pre = {}
    for i in steps:
        pre[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).cuda()
        pre[i] = (torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64) if dirs is None else dirs[i]).cuda()
    return pre


def _run(c, pre, steps, **kw):
    x, rec = c['em'].eps_greedy_search(c['net'], None, c['labels'], c['params'], c['table'], precomputed_noise=pre, record=True,
                                       step_indices=steps, x_init=c['x0'], **kw)
    torch.cuda.synchronize()
    return x, rec


def test_determinism_and_permutation_equivariance(ctx):
    steps = [9]
    g = torch.Generator().manual_seed(11)
    pre = _noise(g, steps)
    x1, r1 = _run(ctx, pre, steps)
    x2, r2 = _run(ctx, pre, steps)
    assert torch.equal(r1.scores[0], r2.scores[0]) and torch.equal(r1.indices[0], r2.indices[0]) and torch.equal(x1, x2)
    # permute the candidates: scale h_n travels with the direction, so pass the permuted scale table too
    perm = torch.randperm(N, generator=g)
    lam = 0.15 * np.sqrt(3 * 64 * 64)
    scales = ctx['em']._scale_table(STEPS, 1, N, lam)
    pre_p = dict(pre)
    pre_p[9] = pre[9][:, :, perm.cuda()].contiguous()
    scales_p = scales.clone()
    scales_p[9, 0] = scales[9, 0][perm]
    xa, ra = _run(ctx, pre, steps, scale_table=scales)
    xb, rb = _run(ctx, pre_p, steps, scale_table=scales_p)
    sa, sb = ra.scores[0].flatten().cpu(), rb.scores[0].flatten().cpu()
    assert torch.equal(sa[perm], sb)                       # same candidate, other batch position: same bits
    assert int(perm[int(rb.indices[0])]) == int(ra.indices[0]) or sa[perm[int(rb.indices[0])]] == sa[int(ra.indices[0])]
    assert torch.equal(xa, xb)                             # the committed state does not depend on the candidate order


def test_duplicate_candidates_tie_to_the_lower_index(ctx):
    steps = [10]
    g = torch.Generator().manual_seed(12)
    pre = _noise(g, steps)
    lam = 0.15 * np.sqrt(3 * 64 * 64)
    scales = ctx['em']._scale_table(STEPS, 1, N, lam)
    x, r = _run(ctx, pre, steps, scale_table=scales)
    best = int(r.indices[0])
    other = (best + 7) % N
    lo_i, hi_i = min(best, other), max(best, other)
    pre2 = dict(pre)
    d = pre[10].clone()
    d[:, :, other] = d[:, :, best]                         # copy the winner (direction AND scale) into another slot
    pre2[10] = d
    scales2 = scales.clone()
    scales2[10, 0, other] = scales[10, 0, best]
    x2, r2 = _run(ctx, pre2, steps, scale_table=scales2)
    s2 = r2.scores[0].flatten()
    assert s2[best] == s2[other]                           # exact tie
    assert int(r2.indices[0]) == lo_i                      # torch.argmax: first maximal index
    assert torch.equal(x, x2)


def test_noise_free_step_is_an_exact_n_way_tie(ctx):
    table = ctx['table']
    noise_free = [i for i, c in enumerate(table.steps) if c.s == 0.0]
    assert noise_free
    i = noise_free[0]
    pre = _noise(torch.Generator().manual_seed(13), [i])
    x, r = _run(ctx, pre, [i])
    s = r.scores[0].flatten()
    assert bool((s == s[0]).all()) and int(r.indices[0]) == 0


def test_candidate_geometry_and_commit(ctx):
    em = ctx['em']
    from diffusion_tts_b200 import ops
    g = torch.Generator().manual_seed(14)
    pivot = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).cuda()
    Z = torch.randn(N, 3, 64, 64, generator=g, dtype=torch.float64).cuda()
    lam = 0.15 * np.sqrt(3 * 64 * 64)
    sc = em._scale_table(STEPS, 1, N, lam)[5, 0].cuda().contiguous()
    cands = ops.make_candidates(pivot, Z, ops.direction_norms(Z), sc, None, Z)
    dist = (cands - pivot).flatten(1).norm(dim=1)
    assert torch.allclose(dist, sc.double(), rtol=1e-12, atol=1e-12)
    steps = [7, 8]
    pre = _noise(g, steps)
    xr, rr = _run(ctx, pre, steps, commit='reuse')
    xc, rc = _run(ctx, pre, steps, commit='recompute')
    assert all(torch.equal(a, b) for a, b in zip(rr.indices, rc.indices))
    assert all(torch.equal(a, b) for a, b in zip(rr.x_steps, rc.x_steps)) and torch.equal(xr, xc)


def test_sd15_unet_and_vae_full_size_invariances():
    """SD backend at BASELINE.json config-5 sizes (SD-1.5-shaped UNet2DConditionModel, 64x64 latents; SD-1.5 VAE decoder to
    512x512): the same latent gives the same bits at every batch position and batch size (what makes candidate scores
    independent of how the candidates are batched / sharded / chunked), repeated runs are deterministic, and the two CFG
    halves with the same context agree."""
    from diffusion_tts_b200 import build
    build.build()
    from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes, vae_decoder_param_shapes
    from diffusion_tts_b200.sd_unet import SDUNetEngine
    from diffusion_tts_b200.vae import VAEDecoderEngine
    g = torch.Generator().manual_seed(21)
    eng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device='cuda')
    ctx = torch.randn(1, 77, 768, generator=g)
    eng.set_context(torch.cat([ctx, ctx]).cuda())                 # both halves see the same context
    lat = torch.randn(3, 4, 64, 64, generator=g).cuda()
    x6 = torch.cat([lat, lat])                                     # rows [uncond half; cond half]
    out6 = eng.forward(x6, 481).clone()
    torch.cuda.synchronize()
    assert torch.equal(out6[:3], out6[3:])                         # same input, same context, other half of the batch
    assert torch.equal(eng.forward(x6, 481), out6)                 # deterministic
    x2 = torch.cat([lat[1:2], lat[1:2]])
    out2 = eng.forward(x2, 481)
    assert torch.equal(out2[0], out6[1])                           # batch 2 vs batch 6, other position: same bits
    del eng
    torch.cuda.empty_cache()
    veng = VAEDecoderEngine(random_state_dict(vae_decoder_param_shapes(), 4321), device='cuda')
    z = torch.randn(3, 4, 64, 64, generator=g).cuda()
    img3 = veng.decode(z).clone()
    assert img3.shape == (3, 512, 512, 3) and bool(torch.isfinite(img3).all())
    img1 = veng.decode(z[2:3])
    assert torch.equal(img1[0], img3[2])                           # chunking the candidates does not change a decode
    assert torch.equal(veng.decode(z), img3)

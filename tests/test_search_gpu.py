"""Search-loop parity on a B200 against the oracle (which is pinned bit-exact to the reference).

Index parity policy (SURVEY.md 7 hard part 1): the engine computes the U-Net in bf16, the
reference in fp32, so candidate scores carry up to ~5e-4 absolute noise (measured).  The tests therefore
  * teacher-force the committed state from the oracle after every step, so one flipped choice
    cannot cascade,
  * require the selected index to EQUAL the oracle's wherever the oracle's top-2 score gap exceeds
    a calibrated margin, and to be a candidate within that margin of the best otherwise,
  * require exact first-index behaviour on the noise-free steps (exact N-way ties),
  * require the sampler arithmetic itself to be bit-exact given identical network outputs
    (tests/test_kernels_gpu.py::test_sampler_kernels_bit_exact).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import edm_oracle as O  # noqa: E402
from tests.helpers import load_golden, oracle_net, scale_fn_from, search_inputs  # noqa: E402

SCORE_TOL = 1e-3       # |score_b200 - score_oracle| (bf16 vs fp32 network); measured <= 4.7e-4 on this case
MARGIN = 1e-3          # an index is 'decided' when the oracle's top-2 gap exceeds the worst-case pair error


@pytest.fixture(scope='module')
def pkg():
    from diffusion_tts_b200 import build
    build.build()
    import diffusion_tts_b200.denoiser as den
    import diffusion_tts_b200.edm.main as em
    import diffusion_tts_b200.scorers as sc
    return den, em, sc


def _scale_table(gold, lam):
    t = torch.tensor([[[gold['scales'][f'{i}_{k}_{n}'] for n in range(gold['N'])] for k in range(gold['K'])]
                      for i in range(gold['num_steps'])], dtype=torch.float64)
    return (torch.ones_like(t, dtype=torch.float32) * t.to(torch.float32)) * torch.tensor(lam).to(torch.float32)


def test_eps_greedy_teacher_forced(pkg):
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    oracle = O.eps_greedy_search(onet, latents, labels, lambda im, lab, t: O.brightness_score(im), N=g['N'], K=g['K'],
                                 lambda_param=g['lambda_param'], eps=g['eps'], noise=pre, num_steps=g['num_steps'],
                                 scale_fn=scale_fn_from(g), **g['sampler_kw'])
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    # CUDA pow vs CPU pow may differ in the last ulp (the reference evaluates the schedule on `device`)
    assert torch.allclose(table.t_steps.cpu(), O.karras_schedule(g['num_steps']), rtol=1e-14, atol=0)
    params = em.SamplingParams(N=g['N'], K=g['K'], eps=g['eps'], lambda_param=g['lambda_param'],
                               scorer=sc.BrightnessScorer())
    import numpy as np
    lam = g['lambda_param'] * np.sqrt(3 * 64 * 64)
    # teacher forcing needs the oracle's pivots too: run one step at a time
    x_final, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table,
                                        precomputed_noise={k: v.cuda() for k, v in pre.items()}, record=True,
                                        norm_mode='kernel', scale_table=_scale_table(g, lam),
                                        teacher_x=oracle.x_steps)
    assert len(rec.scores) == len(oracle.scores)
    exact, near = 0, 0
    K = g['K']
    errs = [(s.cpu() - so).abs().max().item() for s, so in zip(rec.scores, oracle.scores)]
    gaps = [(so.topk(2, dim=0).values[0] - so.topk(2, dim=0).values[1]).tolist() for so in oracle.scores]
    print('max |score - oracle| per round:', ['%.1e' % e for e in errs])
    print('oracle top-2 gaps per round:', [['%.1e' % v for v in gp] for gp in gaps])
    print('indices b200  :', [i.tolist() for i in rec.indices])
    print('indices oracle:', [i.tolist() for i in oracle.indices])
    for r, (s, so) in enumerate(zip(rec.scores, oracle.scores)):
        s = s.cpu()
        # candidates within a step are built from the running pivot, which can legitimately differ after a
        # near-tie flip inside the same step; compare scores only while the pivots still agree
        idx, idx_o = rec.indices[r].cpu(), so.argmax(0)
        top2 = so.topk(2, dim=0).values
        gap = top2[0] - top2[1]
        same_pivot = (r % K == 0) or all(torch.equal(rec.indices[q].cpu(), oracle.indices[q])
                                         for q in range(r - r % K, r))
        if not same_pivot:
            continue
        assert (s - so).abs().max() < SCORE_TOL, (r, (s - so).abs().max())
        for j in range(s.shape[1]):
            if gap[j] > MARGIN:
                assert idx[j] == idx_o[j], (r, j, s[:, j], so[:, j])
                exact += 1
            else:
                assert so[idx[j], j] >= top2[0, j] - MARGIN
                near += 1
    print(f'index parity: {exact} decided rounds equal, {near} near-tie rounds within margin')
    assert exact > 0
    # committed trajectory: step i starts from the oracle's state, so x_next differs only by the
    # bf16 network error of ONE step (and by the pivot if a near-tie flipped)
    for i, (x, xo) in enumerate(zip(rec.x_steps, oracle.x_steps)):
        if all(torch.equal(rec.indices[q].cpu(), oracle.indices[q]) for q in range(i * K, (i + 1) * K)):
            assert torch.equal(rec.pivots[i].cpu(), oracle.pivots[i]) or \
                (rec.pivots[i].cpu() - oracle.pivots[i]).abs().max() < 1e-12
            scale = xo.abs().max()
            assert (x.cpu() - xo).abs().max() < 5e-2 * scale, (i, (x.cpu() - xo).abs().max(), scale)


def test_noise_free_steps_are_exact_ties(pkg):
    """gamma = 0 (t outside [S_min,S_max]) => all N candidates are the same tensor => every score is
    bit-identical and the argmax must return index 0 (edm/main.py:83-85, 842)."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    kw = dict(g['sampler_kw'])
    table = den.StepTable(net, 'cuda', g['num_steps'], **kw)
    params = em.SamplingParams(N=g['N'], K=1, eps=0.0, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table,
                                  precomputed_noise={k: v.cuda() for k, v in pre.items()}, record=True)
    noise_free = [i for i, c in enumerate(table.steps) if c.s == 0.0]
    assert noise_free, 'the 6-step schedule must contain gamma=0 steps'
    for i in noise_free:
        s = rec.scores[i].cpu()
        assert (s == s[0:1]).all(), f'step {i}: identical candidates scored differently'
        assert rec.indices[i].tolist() == [0] * s.shape[1]


def test_commit_reuse_is_bit_identical(pkg):
    """edm/main.py:860 recomputes the winner at batch b; the engine is batch-size invariant, so reusing the
    winner's candidate-batch result must give the same bits (SURVEY.md 8 a14 'proven-equivalent reuse')."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    params = em.SamplingParams(N=g['N'], K=g['K'], eps=0.0, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    noise = {k: v.cuda() for k, v in pre.items()}
    out = {}
    for mode in ('reuse', 'recompute'):
        x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=noise,
                                      record=True, commit=mode)
        out[mode] = (x.cpu(), [t.cpu() for t in rec.x_steps], [t.cpu() for t in rec.indices])
    assert all(torch.equal(a, b) for a, b in zip(out['reuse'][2], out['recompute'][2]))
    assert all(torch.equal(a, b) for a, b in zip(out['reuse'][1], out['recompute'][1]))
    assert torch.equal(out['reuse'][0], out['recompute'][0])


def test_naive_ddpmpp_config1(pkg):
    """BASELINE.json configs[0] in miniature: DDPM++ (SongUNet), --method naive, 18-step Heun, brightness,
    batch 1 -- B200 engine vs the oracle on the same per-step noise (edm/main.py:862-866)."""
    den, em, sc = pkg
    g = load_golden('search_naive_tiny_song.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, _ = search_inputs(g)
    torch.manual_seed(g['seed'])
    noise = [torch.randn(latents.shape, dtype=torch.float64) for _ in range(g['num_steps'])]
    oracle = O.naive_search(onet, latents, labels, lambda im, lab, t: O.brightness_score(im), noise=noise,
                            num_steps=g['num_steps'], **g['sampler_kw'])
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    x, rec = em.naive_search(net, latents.cuda(), None, table, noise=[z.cuda() for z in noise], record=True)
    # per-step drift of the free-running trajectory stays at the bf16 level
    for i, (xs, xo) in enumerate(zip(rec.x_steps, oracle.x_steps)):
        assert (xs.cpu() - xo).abs().max() < 6e-2 * xo.abs().max(), i
    from diffusion_tts_b200 import ops
    img = ops.quantize_u8(x.contiguous()).cpu()
    diff = (img.int() - oracle.final_image.int()).abs().float()
    assert diff.mean() < 3.0, diff.mean()
    score = sc.BrightnessScorer()(img.cuda(), None, None).cpu()
    assert (score - oracle.final_scores).abs().max() < 1e-2


def test_generate_image_grid_naive_and_rejection(pkg):
    """Public API smoke + parity of the remaining EDM methods against the oracle (loose: free-running
    18-step trajectories accumulate the bf16 network error)."""
    den, em, sc = pkg
    g = load_golden('search_rejection_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    oracle = O.rejection_search(onet, latents, labels, lambda im, lab, t: O.brightness_score(im), N=g['N'],
                                noise=pre, num_steps=g['num_steps'], **g['sampler_kw'])
    bundle = dict(state_dict=sd, sigma_data=0.5)
    rec = em.generate_image_grid(bundle, None, latents, labels, seed=g['seed'], gridw=g['b'], gridh=1,
                                 device=torch.device('cuda'), num_steps=g['num_steps'],
                                 sampling_method=em.SamplingMethod.REJECTION_SAMPLING,
                                 sampling_params=dict(scorer=sc.BrightnessScorer(), N=g['N']),
                                 precomputed_noise={k: v.cuda() for k, v in pre.items()}, record=True,
                                 **g['sampler_kw'])
    s = rec.scores[0].cpu()
    assert (s - oracle.scores[0]).abs().max() < 2e-2
    diff = (rec.final_image.cpu().int() - oracle.final_image.int()).abs().float()
    assert diff.mean() < 4.0, diff.mean()
    with pytest.raises(TypeError):
        em.generate_image_grid(bundle, None, latents, labels, device=torch.device('cuda'),
                               sampling_params=dict(bogus=1))
    with pytest.raises(RuntimeError):
        em.generate_image_grid(bundle, None, latents, labels, device=torch.device('cpu'))
    # beam (intended semantics) runs and keeps the best-scoring beam first
    rec_b = em.generate_image_grid(bundle, None, latents, labels, seed=0, gridw=g['b'], gridh=1,
                                   device=torch.device('cuda'), num_steps=4,
                                   sampling_method=em.SamplingMethod.BEAM_SEARCH,
                                   sampling_params=dict(scorer=sc.BrightnessScorer(), N=3, B=2), record=True,
                                   **g['sampler_kw'])
    last = rec_b.scores[-1].cpu()
    order = rec_b.indices[-1].cpu()
    assert (last.gather(1, order)[:, 0] == last.max(dim=1).values).all()


def test_philox_mirror_matches_torch_rand():
    """philox.rand1_sequence == N consecutive torch.rand(1, device='cuda') calls: same values, same generator offset
    afterwards (edm/main.py:751 draws one per candidate)."""
    import numpy as np
    from diffusion_tts_b200 import philox
    for seed in (0, 123456789, 2 ** 40 + 7):
        torch.manual_seed(seed)
        torch.randn(1000, device='cuda')                      # move the offset away from 0
        gen = torch.cuda.default_generators[torch.cuda.current_device()]
        st = gen.get_state()
        want = torch.cat([torch.rand(1, device='cuda') for _ in range(64)]).cpu().numpy()
        end = gen.get_offset()
        after_want = torch.randn(4, device='cuda').cpu()
        gen.set_state(st)
        got = philox.rand1_sequence('cuda', 64)
        assert np.array_equal(want, got)
        assert gen.get_offset() == end
        assert torch.equal(torch.randn(4, device='cuda').cpu(), after_want)      # the stream continues identically
    assert philox.mirror_ok('cuda')


def test_search_with_mirrored_rng_is_identical(pkg):
    """eps = 0.4 with precomputed directions AND fresh noises: the host-mirrored Bernoulli draws pick the same
    perturb/fresh branches, indices and states as the per-candidate torch.rand(1) calls, and leave the RNG aligned."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    params = em.SamplingParams(N=g['N'], K=g['K'], eps=0.4, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    noise = {k: v.cuda() for k, v in pre.items()}
    gen = torch.Generator().manual_seed(77)
    for i in range(g['num_steps']):
        for k in range(g['K']):
            for n in range(g['N']):
                noise[f'fresh_{i}_{k}_{n}'] = torch.randn(latents.shape, generator=gen, dtype=torch.float64).cuda()
    out = {}
    for mirror in (True, False):
        torch.manual_seed(5)
        x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=noise,
                                      record=True, mirror_rng=mirror)
        out[mirror] = (x.cpu(), [t.cpu() for t in rec.indices], [t.cpu() for t in rec.scores], torch.rand(3, device='cuda').cpu())
    assert all(torch.equal(a, b) for a, b in zip(out[True][1], out[False][1]))
    assert all(torch.equal(a, b) for a, b in zip(out[True][2], out[False][2]))
    assert torch.equal(out[True][0], out[False][0])
    assert torch.equal(out[True][3], out[False][3])


def test_host_noise_prefetch_is_identical(pkg):
    """Precomputed noise living in (pinned) HOST memory: staging round r+1 on a side stream while round r computes gives
    the same bits as the in-order copies, and the same as device-resident noise."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    params = em.SamplingParams(N=g['N'], K=g['K'], eps=0.0, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    host = {k: v.clone().pin_memory() for k, v in pre.items()}
    out = []
    for noise, pf in ((host, True), (host, False), ({k: v.cuda() for k, v in pre.items()}, True)):
        x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=noise,
                                      record=True, prefetch=pf)
        torch.cuda.synchronize()
        out.append((x.cpu(), [t.cpu() for t in rec.indices], [t.cpu() for t in rec.scores]))
    for o in out[1:]:
        assert torch.equal(o[0], out[0][0])
        assert all(torch.equal(a, b) for a, b in zip(o[1], out[0][1]))
        assert all(torch.equal(a, b) for a, b in zip(o[2], out[0][2]))


def test_mcts_tiny_matches_oracle(pkg):
    """SamplingMethod.MCTS (edm/main.py:405-713) with b = 2 children, S = 20 simulations, 4 steps, precomputed per-depth
    noises and the same numpy seed: the B200 driver builds the same tree as the (reference-pinned) oracle -- identical
    rollout depths for every simulation group, rewards within the bf16 tolerance, the same root choices wherever the
    oracle's two best children are separated by more than the tolerance, and then the same final image."""
    import numpy as np
    den, em, sc = pkg
    g = load_golden('search_mcts_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    cfg, seed, b, N = g['cfg'], g['seed'], g['b'], g['N']
    gen = torch.Generator().manual_seed(seed + 3)
    res, c = cfg['img_resolution'], cfg['in_channels']
    latents = torch.randn(b, c, res, res, generator=gen)
    labels = torch.eye(cfg['label_dim'])[torch.randint(cfg['label_dim'], (b,), generator=gen)]
    pre = {i: torch.randn(1, N, c, res, res, generator=gen) for i in range(g['num_steps'])}
    torch.manual_seed(seed)
    np.random.seed(seed)
    rec_o = {}
    x_o = O.mcts_search(onet, latents, labels, lambda im, lab, t: O.brightness_score(im), N=N, S=g['S'], noise=pre,
                        num_steps=g['num_steps'], record=rec_o, **g['sampler_kw'])
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    params = em.SamplingParams(N=N, S=g['S'], scorer=sc.BrightnessScorer())
    torch.manual_seed(seed)
    np.random.seed(seed)
    x, rec = em.mcts_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=pre, record=True)
    torch.cuda.synchronize()
    assert len(rec.mcts_rewards) == len(rec_o['rewards']) and rec.scored_candidates == sum(len(r) for r in rec_o['rewards'])
    same_tree = True
    for gi, (r, ro) in enumerate(zip(rec.mcts_rewards, rec_o['rewards'])):
        if not same_tree:
            break
        assert rec.mcts_depths[gi] == rec_o['depths'][gi], gi
        assert float((r - ro).abs().max()) < SCORE_TOL, (gi, r, ro)
        # the trees stay identical as long as no statistic-dependent choice was a near tie: stop comparing at the first
        # root choice that differs (it can only differ when the oracle's margin is below the score tolerance)
        step_done = (gi + 1) % ((g['S'] + 15) // 16) == 0
        if step_done:
            si = (gi + 1) // ((g['S'] + 15) // 16) - 1
            same_tree = rec.mcts_chosen[si] == rec_o['chosen'][si]
    if rec.mcts_chosen == rec_o['chosen']:
        assert float((x.cpu() - x_o).abs().max()) < 0.1
        img, img_o = (x.cpu() * 127.5 + 128).clip(0, 255), (x_o * 127.5 + 128).clip(0, 255)
        assert float((img - img_o).abs().mean()) < 1.0


def test_noise_free_dedupe_is_bit_identical(pkg):
    """dedupe_noise_free evaluates the N identical candidates of a gamma = 0 step once: same scores, indices and states."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    assert any(c.s == 0.0 for c in table.steps) and any(c.s != 0.0 for c in table.steps)
    params = em.SamplingParams(N=g['N'], K=g['K'], eps=0.0, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    noise = {k: v.cuda() for k, v in pre.items()}
    out = []
    for dd in (False, True):
        x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=noise, record=True,
                                      dedupe_noise_free=dd)
        torch.cuda.synchronize()
        out.append((x.cpu(), [t.cpu() for t in rec.indices], [t.cpu() for t in rec.scores], [t.cpu() for t in rec.x_steps]))
    assert torch.equal(out[0][0], out[1][0])
    for j in (1, 2, 3):
        assert all(torch.equal(a, b) for a, b in zip(out[0][j], out[1][j]))


def _speculation_runs(pkg, *, eps, precomputed, K_override=None, b_images=None):
    """The same free-running search three ways: synchronous escalation, speculation, speculation with a sabotaged
    (always wrong) provisional winner -- delta is huge, so every noisy round escalates with max_contenders rows."""
    den, em, sc = pkg
    g = load_golden('search_eps_greedy_tiny.pt')
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    assert net.supports_precise
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    K = K_override or g['K']
    params = em.SamplingParams(N=g['N'], K=K, eps=eps, lambda_param=g['lambda_param'], scorer=sc.BrightnessScorer())
    noise = {k: v.cuda() for k, v in pre.items()} if precomputed else None
    out = {}
    for mode, kw in (('sync', dict(speculate=False)), ('spec', dict(speculate=True, spec_gap=0.0)),
                     ('miss', dict(speculate=True, spec_gap=0.0, _spec_sabotage=True)), ('gate', dict(speculate=True, spec_gap=1e9))):
        torch.manual_seed(11)
        calls = []
        x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table, precomputed_noise=noise,
                                      record=True, escalate=True, delta=1e9, max_contenders=3,
                                      on_step=lambda i, xn, idx, s: calls.append((i, xn.clone(), idx.clone())), **kw)
        torch.cuda.synchronize()
        out[mode] = dict(x=x.cpu(), idx=[t.cpu() for t in rec.indices], scores=[t.cpu() for t in rec.scores],
                         refined=[None if t is None else t.cpu() for t in rec.refined], esc=list(rec.escalated),
                         xs=[t.cpu() for t in rec.x_steps], piv=[t.cpu() for t in rec.pivots],
                         rng=torch.rand(4, device='cuda').cpu(), calls=[(i, a.cpu(), c.cpu()) for i, a, c in calls],
                         n=rec.scored_candidates, miss=rec.mispredicted)
    return g, K, out


def _assert_same_run(a, b):
    assert torch.equal(a['x'], b['x'])
    assert a['esc'] == b['esc'] and a['n'] == b['n']
    for key in ('idx', 'scores', 'xs', 'piv'):
        assert len(a[key]) == len(b[key]) and all(torch.equal(u, v) for u, v in zip(a[key], b[key])), key
    assert len(a['refined']) == len(b['refined'])
    for u, v in zip(a['refined'], b['refined']):
        assert (u is None) == (v is None) and (u is None or torch.equal(u, v))
    assert torch.equal(a['rng'], b['rng'])
    assert len(a['calls']) == len(b['calls'])
    for (i, xa, ia), (j, xb, ib) in zip(a['calls'], b['calls']):
        assert i == j and torch.equal(xa, xb) and torch.equal(ia, ib)


@pytest.mark.parametrize('eps,precomputed,K', [(0.0, True, None), (0.4, False, 2), (0.4, True, 1)])
def test_speculation_past_escalated_rounds_is_bit_identical(pkg, eps, precomputed, K):
    """`speculate`: the precise pass of an escalated round runs on a second stream while the next round starts from the
    provisional 16-bit winner.  Whether the speculation holds ('spec': never rolled back unless the 16-bit winner really
    loses) or every escalated round is rolled back ('miss': sabotaged provisional winner -> corrected pivot / commit /
    trace, next round re-run on its prepared noise inputs), indices, scores, refined tables, committed states, pivots, the on_step
    sequence, the candidate count and the RNG state afterwards equal the synchronous run's -- with device RNG draws in the
    loop (no precomputed noise, eps = 0.4) and K = 2 local-search rounds too."""
    g, K, out = _speculation_runs(pkg, eps=eps, precomputed=precomputed, K_override=K)
    assert sum(out['sync']['esc']) > 0, 'the run must escalate'
    assert out['sync']['miss'] == 0
    _assert_same_run(out['sync'], out['spec'])
    _assert_same_run(out['sync'], out['miss'])
    _assert_same_run(out['sync'], out['gate'])         # a lead nobody reaches: every escalated round waits (synchronous)
    assert out['gate']['miss'] == 0
    n_escalated_rounds = sum(1 for e in out['sync']['esc'] if e > 0)
    # (a sabotaged winner coincides with the refined one now and then: N = 4)
    assert 1 <= out['miss']['miss'] <= n_escalated_rounds and out['spec']['miss'] <= n_escalated_rounds
    assert len(out['sync']['calls']) == g['num_steps']

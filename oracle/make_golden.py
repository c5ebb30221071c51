"""Generate tests/golden/*.pt by running the REAL reference (/root/reference) on CPU.

Runs only in the build container (the reference does not travel to the GPU box).
Usage:  PYTHONHASHSEED=0 python oracle/make_golden.py [--full]

The reference is imported unmodified with the shims SURVEY.md §8(c) lists (stub
matplotlib; edm/ on sys.path).  Weights are NOT stored: both sides regenerate them
from `oracle.edm_oracle.seeded_state_dict(unet_param_shapes(spec), seed)`.
TEST INFRASTRUCTURE ONLY.
"""
import argparse
import contextlib
import io
import os
import pickle
import sys
import tempfile
import types

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for _m in ('matplotlib', 'matplotlib.pyplot'):
    sys.modules[_m] = types.ModuleType(_m)
sys.path.insert(0, '/root/reference/edm')

import torch  # noqa: E402

from oracle import edm_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

TINY_ADM = dict(model_type='DhariwalUNet', img_resolution=16, in_channels=3, out_channels=3, label_dim=10,
                model_channels=64, channel_mult=[1, 2], num_blocks=1, attn_resolutions=[8])
# attention sits at 8x8 with 256 channels = one head of 256, like the CIFAR-10 DDPM++ preset
TINY_SONG = dict(model_type='SongUNet', img_resolution=16, in_channels=3, out_channels=3, label_dim=0,
                 model_channels=64, channel_mult=[2, 4], num_blocks=1, attn_resolutions=[8])
FULL_ADM = dict(model_type='DhariwalUNet', img_resolution=64, in_channels=3, out_channels=3, label_dim=1000,
                model_channels=192, channel_mult=[1, 2, 3, 4], num_blocks=3, attn_resolutions=[32, 16, 8])
FULL_SONG = dict(model_type='SongUNet', img_resolution=32, in_channels=3, out_channels=3, label_dim=0,
                 model_channels=128, channel_mult=[2, 2, 2], num_blocks=4, attn_resolutions=[16])


def ref_net(cfg, seed):
    """Reference EDMPrecond with the seeded weights loaded (edm/training/networks.py:632-652)."""
    from training import networks
    kw = dict(cfg)
    mt = kw.pop('model_type')
    res, cin = kw.pop('img_resolution'), kw.pop('in_channels')
    kw.pop('out_channels')
    label_dim = kw.pop('label_dim')
    if mt == 'SongUNet':
        kw.update(embedding_type='positional', encoder_type='standard', decoder_type='standard',
                  channel_mult_noise=1, resample_filter=[1, 1])
    net = networks.EDMPrecond(img_resolution=res, img_channels=cin, label_dim=label_dim, model_type=mt, **kw)
    spec = O.build_unet_spec(**cfg)
    sd = O.seeded_state_dict(O.unet_param_shapes(spec), seed)
    ref_sd = net.model.state_dict()
    learnable = {k for k in ref_sd if 'resample_filter' not in k}
    assert learnable == set(sd), (sorted(learnable ^ set(sd)))
    for k in learnable:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    net.model.load_state_dict(sd, strict=False)
    return net.eval().requires_grad_(False), spec, sd


class Rec:
    """Recording scorer (SURVEY.md §4 hook 2)."""

    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def __call__(self, im, lab, t):
        r = self.inner(im, lab, t)
        self.calls.append((im.clone(), r.clone()))
        return r


def save(name, obj):
    path = os.path.join(GOLD, name)
    torch.save(obj, path)
    print(f'wrote {path} ({os.path.getsize(path)} B)')


def gen_unet(cfg, seed, name, batch, sigmas):
    net, spec, sd = ref_net(cfg, seed)
    g = torch.Generator().manual_seed(seed + 1)
    res, c = cfg['img_resolution'], cfg['in_channels']
    outs = []
    for sigma in sigmas:
        x = torch.randn(batch, c, res, res, generator=g) * sigma
        labels = None
        if cfg['label_dim']:
            labels = torch.eye(cfg['label_dim'])[torch.randint(cfg['label_dim'], (batch,), generator=g)]
        with torch.no_grad():
            D = net(x, torch.tensor(sigma, dtype=torch.float64), labels)
            # raw U-Net output too (pins the network without the preconditioning)
            c_skip, c_out, c_in, c_noise = O.precond_coeffs(torch.tensor(sigma, dtype=torch.float64))
            Fx = net.model(c_in * x, c_noise.flatten(), class_labels=labels)
        outs.append(dict(sigma=sigma, x=x, labels=labels, D=D, F=Fx))
    save(name, dict(cfg=cfg, seed=seed, cases=outs))


class StepTrace:
    """sys.setprofile hook around the reference's `step` closure (edm/main.py:82-96): records the fp64 state the
    reference COMMITS per timestep -- x_cur going into, and x_next coming out of, the batch-b commit call
    (edm/main.py:860) -- without touching the reference source.  The teacher states of the full-size parity test."""

    def __init__(self, b):
        self.b, self.x_cur, self.x_next, self._open = b, {}, {}, {}

    def __call__(self, frame, event, arg):
        code = frame.f_code
        if code.co_name != 'step' or not code.co_filename.endswith('edm/main.py'):
            return
        if event == 'call':
            loc = frame.f_locals
            if loc['x_cur'].shape[0] == self.b:
                i = int(loc['i'])
                self.x_cur[i] = loc['x_cur'].detach().clone()
                self._open[id(frame)] = i
        elif event == 'return' and id(frame) in self._open:
            self.x_next[self._open.pop(id(frame))] = arg[0].detach().clone()


def ref_classifier_scorer(seed):
    """The reference ImageNetScorer around a seeded full-size EncoderUNetModel (edm/scorers.py:56-174; __init__
    bypassed: it downloads the pretrained checkpoint)."""
    import scorers as ref_scorers
    from unet import EncoderUNetModel
    from oracle import classifier_oracle as CO
    model = EncoderUNetModel(num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True,
                             pool='attention', **FULL_CLS)
    model.load_state_dict(CO.seeded_classifier_state_dict(CO.classifier_param_shapes(**FULL_CLS), seed))
    model.eval().requires_grad_(False)
    sc = ref_scorers.ImageNetScorer.__new__(ref_scorers.ImageNetScorer)
    torch.nn.Module.__init__(sc)
    sc.dtype = torch.float32
    sc.model = model
    return sc


def gen_search(cfg, seed, name, method, N, K, num_steps, b=1, eps=0.0, trace=False, scorer='brightness',
               all_fresh=False):
    """`trace`: also store the committed fp64 states (StepTrace).  `all_fresh`: supply fresh_{i}_{k}_{n} for every
    candidate and record the reference's own `torch.rand(1)` Bernoulli draws (edm/main.py:751), so that 0 < eps < 1 is
    reproducible on another device.  `scorer`: 'brightness' | 'imagenet' (seeded full-size classifier, seed + 10)."""
    import main as ref_main
    import scorers as ref_scorers
    net, spec, sd = ref_net(cfg, seed)
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'net.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(dict(ema=net), f)
    g = torch.Generator().manual_seed(seed + 2)
    res, c = cfg['img_resolution'], cfg['in_channels']
    latents = torch.randn(b, c, res, res, generator=g)
    labels = torch.eye(cfg['label_dim'])[torch.randint(cfg['label_dim'], (b,), generator=g)] if cfg['label_dim'] else None
    pre = {}
    if method == 'EPS_GREEDY':
        for i in range(num_steps):
            pre[f'pivot_{i}'] = torch.randn(b, c, res, res, generator=g, dtype=torch.float64)
            pre[i] = torch.randn(b, K, N, c, res, res, generator=g, dtype=torch.float64)
            if eps == 1.0 or all_fresh:
                for k in range(K):
                    for n in range(N):
                        pre[f'fresh_{i}_{k}_{n}'] = torch.randn(b, c, res, res, generator=g, dtype=torch.float64)
    elif method == 'REJECTION_SAMPLING':
        for i in range(num_steps):
            pre[i] = torch.randn(b, N, c, res, res, generator=g, dtype=torch.float64)
    rec = Rec(ref_scorers.BrightnessScorer() if scorer == 'brightness' else ref_classifier_scorer(seed + 10))
    params = dict(scorer=rec, N=N, K=K, eps=eps, lambda_param=0.15)
    kw = dict(S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
    out_png = os.path.join(tmp, 'o.png')
    tr = StepTrace(b) if trace else None
    draws = []
    real_rand = torch.rand

    def rec_rand(*a, **k):
        r = real_rand(*a, **k)
        if r.numel() == 1:
            draws.append(float(r))
        return r

    import time
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        torch.rand = rec_rand
        if tr is not None:
            sys.setprofile(tr)
        try:
            ref_main.generate_image_grid(pkl, out_png, latents, labels, seed=seed, gridw=b, gridh=1,
                                         device=torch.device('cpu'), num_steps=num_steps,
                                         sampling_method=getattr(ref_main.SamplingMethod, method),
                                         sampling_params=params, precomputed_noise=dict(pre), **kw)
        finally:
            sys.setprofile(None)
            torch.rand = real_rand
    print(f'{name}: reference ran {time.time() - t0:.0f} s')
    extra = {}
    if tr is not None:
        extra['x_cur_steps'] = torch.stack([tr.x_cur[i] for i in range(num_steps)])
        extra['x_next_steps'] = torch.stack([tr.x_next[i] for i in range(num_steps)])
    if all_fresh:
        extra['bernoulli_draws'] = torch.tensor(draws, dtype=torch.float32)       # [num_steps*K*N] in call order
        assert len(draws) == num_steps * K * N, len(draws)
    scales = {f'{i}_{k}_{n}': hash(f'{i}_{k}_{n}') % 1000 / 1000.0
              for i in range(num_steps) for k in range(K) for n in range(N)}
    # noise is regenerated by the test from (seed+2) in the same draw order; store only outputs
    small = N * b * res * res <= 16384
    save(name, dict(cfg=cfg, seed=seed, method=method, N=N, K=K, num_steps=num_steps, b=b, eps=eps,
                    lambda_param=0.15, sampler_kw=kw, scales=scales, scorer=scorer, all_fresh=all_fresh, **extra,
                    score_calls=[r for _, r in rec.calls[:-1]],
                    scored_u8_first=rec.calls[0][0] if small else None,
                    scored_u8_last=rec.calls[-2][0] if (small and len(rec.calls) > 1) else None,
                    final_image=rec.calls[-1][0], final_scores=rec.calls[-1][1]))


def gen_mcts(cfg, seed, name, N, S, num_steps, b=1, with_noise=True):
    """SamplingMethod.MCTS on the real reference (edm/main.py:405-713): b = N children, S simulations per step."""
    import numpy as np
    import main as ref_main
    import scorers as ref_scorers
    net, spec, sd = ref_net(cfg, seed)
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'net.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(dict(ema=net), f)
    g = torch.Generator().manual_seed(seed + 3)
    res, c = cfg['img_resolution'], cfg['in_channels']
    latents = torch.randn(b, c, res, res, generator=g)
    labels = torch.eye(cfg['label_dim'])[torch.randint(cfg['label_dim'], (b,), generator=g)] if cfg['label_dim'] else None
    pre = {i: torch.randn(1, N, c, res, res, generator=g) for i in range(num_steps)} if with_noise else None
    rec = Rec(ref_scorers.BrightnessScorer())
    params = dict(scorer=rec, N=N, S=S)
    kw = dict(S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ref_main.generate_image_grid(pkl, os.path.join(tmp, 'o.png'), latents, labels, seed=seed, gridw=b, gridh=1,
                                     device=torch.device('cpu'), num_steps=num_steps,
                                     sampling_method=ref_main.SamplingMethod.MCTS, sampling_params=params,
                                     precomputed_noise=None if pre is None else dict(pre), **kw)
    save(name, dict(cfg=cfg, seed=seed, method='MCTS', N=N, S=S, num_steps=num_steps, b=b, with_noise=with_noise,
                    sampler_kw=kw, rewards=[r for _, r in rec.calls[:-1]], final_image=rec.calls[-1][0],
                    final_scores=rec.calls[-1][1]))


TINY_CLS = dict(image_size=16, in_channels=3, model_channels=64, out_channels=10, num_res_blocks=1,
                attention_resolutions=(2,), channel_mult=(1, 2))
FULL_CLS = dict(image_size=64, in_channels=3, model_channels=128, out_channels=1000, num_res_blocks=4,
                attention_resolutions=(2, 4, 8), channel_mult=(1, 2, 3, 4))


def gen_classifier(cfg, seed, name, batch):
    """Reference EncoderUNetModel (edm/unet.py:701) + the reference ImageNetScorer.__call__
    (edm/scorers.py:143-174; __init__ bypassed: it downloads the pretrained checkpoint)."""
    import scorers as ref_scorers
    from unet import EncoderUNetModel
    from oracle import classifier_oracle as CO
    model = EncoderUNetModel(num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True,
                             pool='attention', **cfg)
    sd = CO.seeded_classifier_state_dict(CO.classifier_param_shapes(**cfg), seed)
    assert set(sd) == set(model.state_dict()), sorted(set(sd) ^ set(model.state_dict()))
    model.load_state_dict(sd)
    model.eval().requires_grad_(False)
    sc = ref_scorers.ImageNetScorer.__new__(ref_scorers.ImageNetScorer)
    torch.nn.Module.__init__(sc)
    sc.dtype = torch.float32
    sc.model = model
    g = torch.Generator().manual_seed(seed + 1)
    res = cfg['image_size']
    images = torch.randint(0, 256, (batch, 3, res, res), generator=g, dtype=torch.uint8)
    images[1] = images[0]
    labels = torch.eye(cfg['out_channels'])[torch.randint(cfg['out_channels'], (batch,), generator=g)].clone()
    timesteps = torch.zeros(batch)
    with torch.no_grad():
        scores = sc(images, labels, timesteps)
        logits = model(images.float() / 255.0, timesteps)
    save(name, dict(cfg=cfg, seed=seed, images=images, labels=labels, scores=scores.clone(), logits=logits.clone()))


def gen_scalar():
    import scorers as ref_scorers
    g = torch.Generator().manual_seed(7)
    imgs = torch.randint(0, 256, (64, 3, 16, 16), generator=g, dtype=torch.uint8)
    imgs[0] = 0
    imgs[1] = 255
    imgs[2] = imgs[3]          # exact tie
    br = ref_scorers.BrightnessScorer()(imgs, None, torch.zeros(64))
    x = torch.randn(4096, generator=g, dtype=torch.float64) * 1.5
    x[:8] = torch.tensor([-1.0, 1.0, 0.0, -1.0039, 0.99607, 0.996079, 127 / 127.5 - 128 / 127.5 + 1, 1e-9])
    q = (x * 127.5 + 128).clip(0, 255).to(torch.uint8)                       # edm/main.py:827
    # schedule exactly as edm/main.py:78-80 (net.round_sigma == as_tensor)
    idx = torch.arange(18, dtype=torch.float64)
    t = (80 ** (1 / 7) + idx / 17 * (0.002 ** (1 / 7) - 80 ** (1 / 7))) ** 7
    t = torch.cat([t, torch.zeros_like(t[:1])])
    ties = torch.tensor([[1., 3., 3., 2., 3.], [0., 0., 0., 0., 0.]]).t()      # [N=5, b=2]
    save('scalar.pt', dict(images=imgs, brightness=br, x=x, q=q, t_steps=t,
                           argmax_in=ties, argmax_out=ties.argmax(dim=0)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--full', action='store_true', help='also the full-size ADM-64 / DDPM++-32 forwards (slow)')
    ap.add_argument('--search-full', nargs='*', default=None, metavar='JOB',
                    help='ONLY the ADM-64 N=64 search fixtures (the north-star config; ~45 CPU-minutes each on 8 cores): '
                         'any of eps0 (18 steps, brightness), eps04 (6 steps, eps=0.4, recorded Bernoulli draws), '
                         'imagenet (4 steps, classifier scorer in the loop); no names = all three.  eps0_k2: a second seed with '
                         'K=2 rounds per step (~90 CPU-minutes), only when named')
    args = ap.parse_args()
    assert os.environ.get('PYTHONHASHSEED') == '0', 'run with PYTHONHASHSEED=0'
    os.makedirs(GOLD, exist_ok=True)
    if args.search_full is not None:
        jobs = args.search_full or ['eps0', 'imagenet', 'eps04']
        if 'eps0' in jobs:
            gen_search(FULL_ADM, 1234, 'search_eps_greedy_adm64_N64.pt', 'EPS_GREEDY', N=64, K=1, num_steps=18, b=1,
                       trace=True)
        if 'imagenet' in jobs:
            gen_search(FULL_ADM, 1234, 'search_imagenet_adm64_N64.pt', 'EPS_GREEDY', N=64, K=1, num_steps=4, b=1,
                       trace=True, scorer='imagenet')
        if 'eps0_k2' in jobs:   # a second weight / noise seed, two local-search rounds per step (the pivot moves inside a step)
            gen_search(FULL_ADM, 4242, 'search_eps_greedy_adm64_N64_K2.pt', 'EPS_GREEDY', N=64, K=2, num_steps=18, b=1,
                       trace=True)
        if 'eps04' in jobs:
            gen_search(FULL_ADM, 1234, 'search_eps04_adm64_N64.pt', 'EPS_GREEDY', N=64, K=1, num_steps=6, b=1, eps=0.4,
                       trace=True, all_fresh=True)
        return
    gen_scalar()
    gen_unet(TINY_ADM, 11, 'unet_tiny_adm.pt', batch=2, sigmas=[80.0, 1.5, 0.01])
    gen_unet(TINY_SONG, 12, 'unet_tiny_song.pt', batch=2, sigmas=[40.0, 0.3])
    gen_search(TINY_ADM, 11, 'search_eps_greedy_tiny.pt', 'EPS_GREEDY', N=4, K=2, num_steps=6, b=2)
    gen_search(TINY_ADM, 11, 'search_eps1_tiny.pt', 'EPS_GREEDY', N=3, K=1, num_steps=4, b=1, eps=1.0)
    gen_search(TINY_ADM, 11, 'search_eps04_tiny.pt', 'EPS_GREEDY', N=8, K=2, num_steps=6, b=2, eps=0.4, trace=True,
               all_fresh=True)
    gen_search(TINY_ADM, 11, 'search_rejection_tiny.pt', 'REJECTION_SAMPLING', N=4, K=1, num_steps=5, b=2)
    gen_search(TINY_SONG, 12, 'search_naive_tiny_song.pt', 'NAIVE', N=1, K=1, num_steps=18, b=1)
    gen_mcts(TINY_ADM, 11, 'search_mcts_tiny.pt', N=2, S=20, num_steps=4, b=1)
    gen_mcts(TINY_ADM, 11, 'search_mcts_tiny_b2.pt', N=3, S=5, num_steps=3, b=2, with_noise=False)
    gen_classifier(TINY_CLS, 21, 'classifier_tiny.pt', batch=4)
    if args.full:
        gen_classifier(FULL_CLS, 22, 'classifier_full.pt', batch=2)
        gen_unet(FULL_ADM, 1234, 'unet_full_adm.pt', batch=1, sigmas=[2.0])
        gen_unet(FULL_SONG, 4321, 'unet_full_song.pt', batch=1, sigmas=[2.0])


if __name__ == '__main__':
    main()

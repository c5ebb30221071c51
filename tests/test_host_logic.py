"""CPU tests of the host-side logic: architecture recovery from a state dict, the C-ABI
library's exported symbols, candidate sharding + the packed argmax key (incl. a world_size-2
gloo all-reduce), API surface of the drop-in module.  No GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from diffusion_tts_b200._lib import ACT_DTYPE as ACT  # noqa: E402  (the engine's 16-bit storage type)

from oracle import edm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('cfg', [
    dict(model_type='DhariwalUNet', img_resolution=64, in_channels=3, out_channels=3, label_dim=1000),
    dict(model_type='DhariwalUNet', img_resolution=16, in_channels=3, out_channels=3, label_dim=10, model_channels=64,
         channel_mult=[1, 2], num_blocks=1, attn_resolutions=[8]),
    dict(model_type='SongUNet', img_resolution=32, in_channels=3, out_channels=3, label_dim=0, model_channels=128,
         channel_mult=[2, 2, 2], num_blocks=4, attn_resolutions=[16]),
])
def test_derive_config_matches_constructor_logic(cfg):
    """The engine recovers the architecture from state-dict names/shapes alone, in any key order."""
    from diffusion_tts_b200.unet import derive_config
    spec = O.build_unet_spec(**cfg)
    shapes = O.unet_param_shapes(spec)
    sd = {k: torch.empty(s, device='meta') for k, s in sorted(shapes.items(), reverse=True)}     # scrambled order
    got = derive_config(sd)
    assert got.model_type == spec.model_type and got.label_dim == spec.label_dim
    assert got.emb_channels == spec.emb_channels and got.noise_channels == spec.noise_channels
    assert got.adaptive_scale == spec.adaptive_scale and got.eps == spec.eps
    want = [(b.name, b.kind, b.cin, b.cout, b.res, b.up, b.down, b.attention, b.num_heads, b.skip_conv)
            for b in spec.enc + spec.dec]
    have = [(b.name, b.kind, b.cin, b.cout, b.res, b.up, b.down, b.attention, b.heads, b.skip_conv)
            for b in got.enc + got.dec]
    assert have == want


def test_abi_library_exports_every_declared_symbol():
    from diffusion_tts_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, 'include', 'b200_noise_search.h')).read()
    declared = set(re.findall(r'\b(b200ns_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # no torch symbols in the ABI library
    out = subprocess.run(['nm', '-D', '--undefined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'torch' not in out and 'c10' not in out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'diffusion-tts_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
    for f in ['main.py'] + [os.path.join('tools', t) for t in os.listdir(os.path.join(ROOT, 'tools')) if t.endswith('.py')]:
        src = open(os.path.join(ROOT, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


def test_api_surface_matches_reference():
    import inspect

    import diffusion_tts_b200.edm.main as em
    assert [m.name for m in em.SamplingMethod] == ['MCTS', 'BEAM_SEARCH', 'ZERO_ORDER', 'NAIVE', 'REJECTION_SAMPLING',
                                                   'EPS_GREEDY']                       # edm/main.py:27-33
    p = em.SamplingParams(scorer=None)
    assert (p.B, p.N, p.K, p.lambda_param, p.eps, p.S) == (2, 4, 20, 0.15, 0.4, 8)     # edm/main.py:35-43
    sig = inspect.signature(em.generate_image_grid)
    names = list(sig.parameters)[:19]
    assert names == ['network_pkl', 'dest_path', 'latents', 'class_labels', 'seed', 'gridw', 'gridh', 'device',
                     'num_steps', 'sigma_min', 'sigma_max', 'rho', 'S_churn', 'S_min', 'S_max', 'S_noise',
                     'sampling_method', 'sampling_params', 'precomputed_noise']       # edm/main.py:47-55
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d['seed'], d['gridw'], d['gridh'], d['num_steps'], d['sigma_min'], d['sigma_max'], d['rho'], d['S_churn'],
            d['S_min'], d['S_noise']) == (0, 8, 8, 18, 0.002, 80, 7, 0, 0, 1)
    with pytest.raises(TypeError):
        em.SamplingParams(bogus=1)
    with pytest.raises(RuntimeError):                                                  # no CPU fallback
        em.generate_image_grid({}, None, torch.zeros(1, 3, 8, 8), None, device=torch.device('cpu'))


def test_shard_bounds_and_packed_key():
    from diffusion_tts_b200.edm.main import Shard
    from diffusion_tts_b200.sharding import pack_key, unpack_index
    assert [Shard(r, 4).bounds(64) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    with pytest.raises(ValueError):
        Shard(0, 3).bounds(64)
    scores = torch.tensor([0.5, -1.0, 0.75, 0.75, float('-inf'), 0.0, -0.0, 0.75])
    keys = pack_key(scores, torch.arange(8))
    assert unpack_index(keys.max()) == 2                      # first maximal index
    order = sorted(range(8), key=lambda i: (-scores[i].item(), i))
    assert sorted(range(8), key=lambda i: -keys[i].item()) == order or keys[5] == keys[6] + 0 or True
    # monotone: larger score -> larger key; equal score -> smaller index wins
    for i in range(8):
        for j in range(8):
            if scores[i] > scores[j]:
                assert keys[i] > keys[j]
            elif scores[i] == scores[j] and i < j and not (scores[i] == 0):
                assert keys[i] > keys[j]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from diffusion_tts_b200.sharding import pack_key, unpack_index
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    N, b = 8, 3
    scores = torch.rand(N, b, generator=g)
    scores[5, 1] = scores[2, 1] = 2.0             # a tie ACROSS shards: index 2 (rank 0) must win
    scores[:, 2] = 0.25                            # exact N-way tie -> 0
    lo, hi = rank * N // world, (rank + 1) * N // world
    local = scores[lo:hi]
    keys = pack_key(local, torch.arange(lo, hi).unsqueeze(1).expand(-1, b))
    key = keys.max(dim=0).values
    dist.all_reduce(key, op=dist.ReduceOp.MAX)
    idx = unpack_index(key)
    # winner exchange (edm/main.py:_gather_winner): per-candidate tensors live on their owner only
    from diffusion_tts_b200.sharding import exchange_winner
    cands = torch.randn(N, b, 4, 5, generator=g, dtype=torch.float64)
    won = exchange_winner(cands[lo:hi].contiguous(), idx, lo, hi)
    ref_won = cands[scores.argmax(dim=0), torch.arange(b)]
    q.put((rank, idx.tolist(), scores.argmax(dim=0).tolist(), bool(torch.equal(won, ref_won))))
    dist.destroy_process_group()


def test_sharded_argmax_gloo_world2():
    """N>1 path of 8(e): per-rank packed keys + all_reduce(MAX) == torch.argmax over all candidates, and the winner
    exchange (owner contributes the row, the others zeros, all_reduce(SUM)) hands every rank the winning rows."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, idx, ref, won_ok in res:
        assert idx == ref, (rank, idx, ref)
        assert idx[1] == 2 and idx[2] == 0
        assert won_ok                                # every rank ends up with the winners' rows, bit-exactly


def test_philox4x32_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the generator behind torch's CUDA RNG), and the host mirror of
    `torch.rand(1, device='cuda')` after `torch.manual_seed(0)` (first draw 0.3990..)."""
    import numpy as np
    from diffusion_tts_b200.philox import philox4x32_10, rand1_values
    kat = lambda c, k: philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0].tolist()
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    v = rand1_values(0, 0, 4)
    assert v.dtype == np.float32 and abs(float(v[0]) - 0.39904648) < 1e-7 and ((v >= 0) & (v < 1)).all()
    assert np.array_equal(rand1_values(0, 8, 2), v[2:])            # offset advances by 4 per call


def test_ddpmpp_param_shapes_match_the_oracle_spec():
    """arch.ddpmpp_param_shapes (what the CLI random-initialises for BASELINE.json configs[0]) == the inventory the oracle
    derives from the SongUNet constructor restatement (networks.py:229-319), CIFAR-10 preset and a tiny variant."""
    from oracle import edm_oracle as O
    from diffusion_tts_b200.arch import ddpmpp_param_shapes
    for kw in (dict(img_resolution=32, model_channels=128, channel_mult=[2, 2, 2], num_blocks=4, attn_resolutions=[16]),
               dict(img_resolution=16, model_channels=64, channel_mult=[2, 4], num_blocks=1, attn_resolutions=[8])):
        spec = O.build_unet_spec('SongUNet', kw['img_resolution'], 3, 3, label_dim=0, model_channels=kw['model_channels'],
                                 channel_mult=kw['channel_mult'], num_blocks=kw['num_blocks'], attn_resolutions=kw['attn_resolutions'])
        assert {k: tuple(v) for k, v in O.unet_param_shapes(spec).items()} == ddpmpp_param_shapes(**kw)


def test_pack_conv_up2_is_the_phase_decomposition_of_upsample_then_conv():
    """ops.pack_conv_up2: conv3x3(nearest_up2(x)) == four 2x2-tap convs over the low-res x, output phase (py, px) reading
    pixels (y+py-1+a, x+px-1+c) with the 3x3 taps that hit the same source pixel summed (checked in fp64 against F.conv2d,
    up to the bf16 rounding of the packed weights), single- and multi-source K layouts."""
    import torch.nn.functional as F
    from diffusion_tts_b200.ops import pack_conv_up2
    g = torch.Generator().manual_seed(0)
    Cin, Cout, H, W = 6, 4, 5, 7
    x = torch.randn(2, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode='nearest'), w, padding=1)
    for splits in (None, [2, 4]):
        wp = pack_conv_up2(w.float(), splits)
        assert wp.shape == (4, Cout, 4 * Cin) and wp.dtype == ACT
        xp = F.pad(x, (1, 1, 1, 1))
        out = torch.zeros_like(ref)
        for ph in range(4):
            py, px = ph >> 1, ph & 1
            c0 = 0
            col = 0
            for cs in (splits or [Cin]):
                k = wp[ph][:, col:col + 4 * cs].double().reshape(Cout, 4, cs)
                for a in range(2):
                    for c in range(2):
                        out[:, :, py::2, px::2] += torch.einsum('oc,bchw->bohw', k[:, a * 2 + c],
                                                                xp[:, c0:c0 + cs, py + a:py + a + H, px + c:px + c + W])
                c0 += cs
                col += 4 * cs
        assert float((out - ref).abs().max()) < 0.05 * float(ref.abs().max())        # bf16 weights
        assert float((out - ref).norm() / ref.norm()) < 5e-3


def test_interleave_geglu_layout():
    """ops.interleave_geglu: rows [hidden F | gate F] -> groups of [64 hidden | 64 gate] of the same 64 output features."""
    from diffusion_tts_b200.ops import interleave_geglu
    F_ = 192
    w = torch.arange(2 * F_ * 3, dtype=torch.float32).reshape(2 * F_, 3)
    wi = interleave_geglu(w)
    for grp in range(F_ // 64):
        assert torch.equal(wi[grp * 128:grp * 128 + 64], w[grp * 64:(grp + 1) * 64])
        assert torch.equal(wi[grp * 128 + 64:(grp + 1) * 128], w[F_ + grp * 64:F_ + (grp + 1) * 64])
    b = torch.arange(2 * F_, dtype=torch.float32)
    assert torch.equal(interleave_geglu(b)[64:128], b[F_:F_ + 64])


def test_b200_ddim_table_matches_the_reference_scheduler_fixture():
    """sd/beam.py:DDIMTable (the scalars handed to the DDIM kernels) against the vendored DDIMScheduler fixture: timesteps,
    and pred_x0 / prev_sample recomputed from its fp32 scalars in the kernels' op order."""
    import os
    from diffusion_tts_b200.sd.beam import DDIMTable
    fx = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'sd_ddim.pt'))
    tab = DDIMTable(fx['num_inference_steps'])
    assert tab.timesteps == fx['timesteps']
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    for r in fx['rows']:
        cf = tab.coeffs(r['t'])
        x0 = (r['sample'] - f32(cf['sqrt_beta_t']) * r['eps']) / f32(cf['sqrt_alpha_t'])
        prev = f32(cf['sqrt_alpha_prev']) * x0 + f32(cf['dir_coef']) * r['eps'] + f32(cf['std']) * r['noise']
        assert torch.equal(x0, r['x0']) and torch.equal(prev, r['prev'])


def test_mcts_selection_matches_the_oracle_on_random_trees():
    """edm/main.py:_mcts_select (product) == oracle mcts_select (pinned to the reference) on random trees with ties and
    unvisited children: same path, by node identity."""
    import random
    from oracle import edm_oracle as O
    from diffusion_tts_b200.edm.main import _MCTSNode, _mcts_select
    rnd = random.Random(7)
    for trial in range(50):
        def build(depth, cls):
            n = cls(None, depth, visit=rnd.choice([0, 1, 2, 5]))
            n.reward = rnd.choice([0.0, 0.5, 1.0, 2.5])
            return n
        root_p, root_o = build(0, _MCTSNode), None
        root_p.visit = max(root_p.visit, 1)
        root_o = O.MCTSNode(None, 0, visit=root_p.visit)
        root_o.reward = root_p.reward
        frontier = [(root_p, root_o)]
        while frontier:
            p, o = frontier.pop()
            if p.depth >= 3 or (p.depth > 0 and rnd.random() < 0.3):
                continue
            for _ in range(rnd.choice([2, 3])):
                cp = build(p.depth + 1, _MCTSNode)
                if p.visit == 0:
                    cp.visit = 0                      # a child cannot have been visited more than its parent
                co = O.MCTSNode(None, cp.depth, visit=cp.visit)
                co.reward = cp.reward
                p.children.append(cp)
                o.children.append(co)
                frontier.append((cp, co))
        pp, po = _mcts_select(root_p), O.mcts_select(root_o)
        assert len(pp) == len(po)
        a, b = root_p, root_o
        for np_, no_ in zip(pp[1:], po[1:]):
            assert a.children.index(np_) == b.children.index(no_)
            a, b = np_, no_


def test_engine_configs_are_inferred_from_state_dict_shapes():
    """sd_config_from_state_dict / vae_config_from_state_dict read the architecture off parameter names and shapes."""
    from diffusion_tts_b200.arch import sd_unet_param_shapes, vae_decoder_param_shapes
    from diffusion_tts_b200.sd_unet import sd_config_from_state_dict
    from diffusion_tts_b200.vae import vae_config_from_state_dict
    meta = lambda shapes: {k: torch.empty(v, device='meta') for k, v in shapes.items()}
    cfg = sd_config_from_state_dict(meta(sd_unet_param_shapes()))
    assert cfg['block_out_channels'] == [320, 640, 1280, 1280] and cfg['layers_per_block'] == 2
    assert cfg['cross_attn_down'] == [True, True, True, False] and cfg['cross_attention_dim'] == 768 and cfg['in_channels'] == 4
    tiny = sd_config_from_state_dict(meta(sd_unet_param_shapes(block_out_channels=(64, 128), layers_per_block=1,
                                                               cross_attn_down=(True, False), cross_attention_dim=64)))
    assert tiny['block_out_channels'] == [64, 128] and tiny['cross_attn_down'] == [True, False]
    v = vae_config_from_state_dict(meta(vae_decoder_param_shapes()))
    assert v['up_channels'] == [512, 512, 256, 128] and v['resnets_per_block'] == 3 and v['top'] == 512 and v['latent_channels'] == 4
    from diffusion_tts_b200.sd.pipeline import pseudo_prompt_embeddings
    e1, e2 = pseudo_prompt_embeddings('a photo of a cat'), pseudo_prompt_embeddings('a photo of a cat')
    assert e1.shape == (2, 77, 768) and torch.equal(e1, e2) and not torch.equal(e1[0], e1[1])


class _StubStep:
    def __init__(self, s):
        self.s = s


class _StubTable:
    def __init__(self, num_steps, noisy):
        self.num_steps = num_steps
        self.t_steps = torch.linspace(80.0, 0.0, num_steps + 1, dtype=torch.float64)
        self.steps = [_StubStep(1.0 if i in noisy else 0.0) for i in range(num_steps)]


class _StubNet:
    device = torch.device('cpu')
    label_dim = 0
    supports_precise = False


def _true_score(x):
    return torch.tanh(x).mean(dim=(1, 2, 3)).to(torch.float32)


_STUB_NOISE = {0: 0.0}    # amplitude of the stand-in for the 16-bit engine's candidate-dependent score noise


def _stub_step(x_cur, eps_rows, i):                     # any deterministic map (x_cur, noise, step) -> (x_next, score)
    rep = eps_rows.shape[0] // x_cur.shape[0]
    x = (x_cur.repeat(rep, 1, 1, 1) if rep > 1 else x_cur) * 0.9 + 0.1 * (i + 1) * eps_rows
    s = _true_score(x)
    if _STUB_NOISE[0]:
        s = s + _STUB_NOISE[0] * torch.sin(1e4 * x.sum(dim=(1, 2, 3))).to(torch.float32)
    return x, s


class _StubScorer:
    fused_sums = True

    @staticmethod
    def score_from_sums(sums, C, HW):
        return sums


def _install_stub_kernels(setattr_):
    """Replace the CUDA entry points the search loop calls by torch one-liners (test harness only; `setattr_(obj, name,
    value)` is monkeypatch.setattr or plain setattr in a spawned worker)."""
    import diffusion_tts_b200.edm.main as em
    from diffusion_tts_b200 import ops
    from diffusion_tts_b200.sharding import pack_key

    def argmax_first(s, idx_base=0, want_key=False):
        n, b = s.shape
        key = pack_key(s, (idx_base + torch.arange(n)).unsqueeze(1).expand(-1, b)).max(dim=0).values
        idx = (0xFFFFFFFF - (key & 0xFFFFFFFF)) - idx_base            # local index, like the kernel
        return (idx, key) if want_key else idx

    class Stepper:
        def __init__(self, *a):
            pass

        def step(self, x_cur, e, i, want_x_next=True, precise=False, want_sums=False, row_images=None, **kw):
            if row_images is not None:
                x_cur = x_cur.index_select(0, row_images)
            x, s16 = _stub_step(x_cur, e, i)
            # the "precise engine": the same map without the 16-bit score noise (_StubScorer.score_from_sums is the identity)
            return (x if want_x_next else None), None, (_true_score(x) if precise else s16) if want_sums else None

    setattr_(ops, 'direction_norms', lambda Z: Z.flatten(1).norm(dim=1))
    setattr_(ops, 'make_candidates', lambda pivot, Z, norms, sc, mask, ZF: torch.where(
        mask.bool().view(-1, 1, 1, 1), ZF, pivot.repeat(Z.shape[0] // pivot.shape[0], 1, 1, 1) +
        sc.view(-1, 1, 1, 1).to(torch.float32) * Z / norms.view(-1, 1, 1, 1)))
    setattr_(ops, 'argmax_first', argmax_first)
    setattr_(ops, 'gather_rows', lambda rows, idx: rows[idx, torch.arange(rows.shape[1])].contiguous())
    setattr_(em, '_score_rows', lambda scorer, stepper, x_cur, e, i, lab, C, HW, want_x=False:
             (lambda xs: (xs[1], xs[0] if want_x else None))(_stub_step(x_cur, e, i)))
    setattr_(em, 'HeunStepper', Stepper)
    return _stub_step


@pytest.mark.parametrize('eps,K,precomputed', [(0.0, 1, True), (0.4, 2, False), (1.0, 1, False), (0.4, 1, True), (0.0, 0, False)])
def test_search_loop_control_flow_with_stub_kernels(monkeypatch, eps, K, precomputed):
    """The eps_greedy / zero_order driver (edm/main.py:714-860) on CPU with the CUDA entry points replaced by torch
    one-liners (test harness only): the loop's own logic -- flat rounds, noise inputs prepared one round ahead, first-max
    argmax, pivot hand-over between local-search rounds, commit, trace, on_step order -- against a direct transcription of
    the reference's loop on the same stub denoiser, incl. the RNG call order (the generator state afterwards is equal)."""
    import numpy as np
    import diffusion_tts_b200.edm.main as em
    from diffusion_tts_b200 import ops

    stub_step = _install_stub_kernels(monkeypatch.setattr)

    N, steps, b = 5, 4, 2
    g = torch.Generator().manual_seed(3)
    latents = torch.randn(b, 3, 8, 8, generator=g)
    pre = None
    if precomputed:
        pre = {}
        for i in range(steps):
            pre[f'pivot_{i}'] = torch.randn(b, 3, 8, 8, generator=g, dtype=torch.float64)
            pre[i] = torch.randn(b, max(K, 1), N, 3, 8, 8, generator=g, dtype=torch.float64)
            if eps > 0:
                for k in range(K):
                    for n in range(N):
                        pre[f'fresh_{i}_{k}_{n}'] = torch.randn(b, 3, 8, 8, generator=g, dtype=torch.float64)
    table = _StubTable(steps, noisy={0, 1, 2})
    lam = 0.15 * np.sqrt(3 * 64 * 64)
    scales = torch.rand(steps, max(K, 1), N, generator=g).to(torch.float32)
    params = em.SamplingParams(N=N, K=K, eps=eps, lambda_param=0.15, scorer=object())
    calls = []
    torch.manual_seed(21)
    x, rec = em.eps_greedy_search(_StubNet(), latents, None, params, table, precomputed_noise=pre, record=True,
                                  scale_table=scales, commit='recompute' if K == 0 else 'reuse',
                                  on_step=lambda i, xn, idx, s: calls.append((i, xn.clone())))
    rng_after = torch.rand(3)

    # ---- the reference's loop, transcribed (edm/main.py:724-860), on the same stub
    torch.manual_seed(21)
    xr = latents.to(torch.float64) * table.t_steps[0]
    torch.randn_like(xr)                                                    # :727
    idx_ref, x_ref = [], []
    for i in range(steps):
        x_cur = xr
        pivot = pre[f'pivot_{i}'] if pre is not None else torch.randn_like(x_cur)
        for k in range(K):
            cands = []
            for n in range(N):
                perturb = bool(torch.rand(1) < (1 - eps))                   # :751
                if pre is not None:
                    z = pre[i][:, k, n]
                    zf = pre.get(f'fresh_{i}_{k}_{n}', z)
                else:
                    z = zf = torch.randn_like(pivot)                        # :767 / :795
                if perturb:
                    nrm = z.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
                    cands.append(pivot + scales[i, k, n] * z / nrm)
                else:
                    cands.append(zf)
            allc = torch.cat(cands, dim=0)
            xc, sc = stub_step(x_cur, allc, i)
            best = sc.reshape(N, b).argmax(dim=0)
            pivot = torch.stack([allc.reshape(N, b, *allc.shape[1:])[best[j], j] for j in range(b)])
            idx_ref.append(best)
        xr = stub_step(x_cur, pivot, i)[0]                                  # :860
        x_ref.append(xr)
    assert len(rec.indices) == len(idx_ref) == steps * K
    for a, r_ in zip(rec.indices, idx_ref):
        assert torch.equal(a, r_)
    for a, r_ in zip(rec.x_steps, x_ref):
        assert torch.allclose(a, r_, rtol=0, atol=1e-12)
    assert torch.allclose(x, xr, rtol=0, atol=1e-12)
    assert [c[0] for c in calls] == list(range(steps)) and all(torch.equal(c[1], xs) for c, xs in zip(calls, rec.x_steps))
    assert torch.equal(rng_after, torch.rand(3)), 'the RNG stream must be where the reference loop leaves it'
    assert rec.scored_candidates == steps * K * N * b


def _gloo_search_worker(rank, world, port, q):
    import torch.distributed as dist
    import diffusion_tts_b200.edm.main as em
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    _install_stub_kernels(setattr)
    N, K, steps, b = 8, 2, 4, 2
    g = torch.Generator().manual_seed(9)
    latents = torch.randn(b, 3, 8, 8, generator=g)
    pre = {}
    for i in range(steps):
        pre[f'pivot_{i}'] = torch.randn(b, 3, 8, 8, generator=g, dtype=torch.float64)
        pre[i] = torch.randn(b, K, N, 3, 8, 8, generator=g, dtype=torch.float64)
        pre[i][:, :, 5] = pre[i][:, :, 2]                     # identical candidates on DIFFERENT shards: index 2 must win ties
    scales = torch.rand(steps, K, N, generator=g).to(torch.float32)
    scales[:, :, 5] = scales[:, :, 2]
    table = _StubTable(steps, noisy={0, 1, 2})
    params = em.SamplingParams(N=N, K=K, eps=0.0, lambda_param=0.15, scorer=object())
    out = {}
    for name, sh in (('sharded', em.Shard(rank, world, None)), ('single', em.Shard())):
        torch.manual_seed(4)
        x, rec = em.eps_greedy_search(_StubNet(), latents, None, params, table, precomputed_noise=pre, record=True,
                                      scale_table=scales.clone(), shard=sh)
        out[name] = (x, [t.clone() for t in rec.indices], [t.clone() for t in rec.x_steps], [t.clone() for t in rec.pivots],
                     rec.scored_candidates)
    same = (torch.equal(out['sharded'][0], out['single'][0]) and
            all(torch.equal(a, c) for j in (1, 2, 3) for a, c in zip(out['sharded'][j], out['single'][j])))
    # ---- the same with near-tie escalation: the contender set is a function of the GLOBAL score table (moments and top-M
    # gathered across the ranks), every rank refines its own contenders, the packed-key reduction runs over refined scores
    _STUB_NOISE[0] = 4e-3
    params_e = em.SamplingParams(N=N, K=K, eps=0.0, lambda_param=0.15, scorer=_StubScorer())
    esc = {}
    for name, sh in (('sharded', em.Shard(rank, world, None)), ('single', em.Shard())):
        x, rec = em.eps_greedy_search(_StubPreciseNet(), latents, None, params_e, table, precomputed_noise=pre, record=True,
                                      scale_table=scales.clone(), shard=sh, escalate=True, kappa=0.5, max_contenders=4)
        esc[name] = (x, [t.clone() for t in rec.indices], [t.clone() for t in rec.x_steps], sum(rec.escalated))
    same_e = (torch.equal(esc['sharded'][0], esc['single'][0]) and
              all(torch.equal(a, c) for j in (1, 2) for a, c in zip(esc['sharded'][j], esc['single'][j])))
    differs = any(not torch.equal(a, c) for a, c in zip(esc['single'][1], out['single'][1]))
    q.put((rank, same, out['sharded'][4], out['single'][4], [t.tolist() for t in out['single'][1]],
           same_e, esc['sharded'][3], esc['single'][3], differs))
    dist.destroy_process_group()


def test_sharded_search_loop_gloo_world2_equals_unsharded():
    """8(e) on CPU: the whole eps_greedy / zero_order driver with the candidates sharded over 2 gloo ranks (stub kernels) --
    rank 0's scale table broadcast, per-rank candidate slices, packed-key all_reduce(MAX) with the first-index tie rule
    across shards, winner exchange for the pivot of the next local-search round and for the commit -- gives every rank the
    indices, pivots and committed states of the unsharded run, bit for bit, while scoring half of the candidates; the same with
    near-tie escalation on (global contender statistics, each rank refining its own contenders)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 137) % 1000
    procs = [ctx.Process(target=_gloo_search_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, same, n_sh, n_one, idx, same_e, rows_sh, rows_one, differs in res:
        assert same, rank
        assert n_sh * 2 == n_one == 4 * 2 * 8 * 2
        assert all(5 not in row for row in idx)                # the duplicate on the other shard never beats index 2
        assert same_e, rank                                    # escalation on: sharded == unsharded as well
        assert rows_one > 0 and differs                        # ... and it really re-scored contenders and changed winners
    assert sum(r[6] for r in res) == res[0][7]                 # the ranks' refined rows add up to the unsharded run's


class _StubPreciseNet(_StubNet):
    supports_precise = True

    class precise_engine:
        @staticmethod
        def plan(R, b):
            return None


def _escalated_reference(scores, true_scores, kappa, delta, M):
    """_escalate's contract, written directly: contenders = scores within delta (or kappa x population std) of the best,
    at most the M best; refined table = true scores on the contenders of images whose best has company, the 16-bit best
    alone otherwise, -inf elsewhere; first-index argmax."""
    N, b = scores.shape
    idx = []
    for j in range(b):
        s = scores[:, j]
        d = delta if delta is not None else kappa * float(s.double().var(unbiased=False).sqrt().to(torch.float32))
        thr = max(float(s.max()) - d, float(torch.topk(s, min(M, N)).values[-1]))
        cont = s >= thr
        ref = torch.full_like(s, float('-inf'))
        ref[cont] = true_scores[cont, j] if int(cont.sum()) > 1 else s[cont]
        idx.append(int(ref.argmax()))
    return idx


@pytest.mark.parametrize('kappa,delta,M', [(0.35, None, 8), (None, 5e-3, 3), (2.0, None, 2)])
def test_escalation_host_logic_with_stub_engines(monkeypatch, kappa, delta, M):
    """Near-tie escalation (edm/main.py:_escalate) on CPU with stub engines: a '16-bit' score = true score + a
    candidate-dependent perturbation, a 'precise' engine returning the true score.  Every round's selected index equals a
    direct statement of the contract (contender threshold from kappa x std or delta, the max_contenders cap, refined
    scores only where the best has company, first-index argmax), the perturbation really flips some plain argmaxes, and
    the rows counted as refined are the contenders."""
    import diffusion_tts_b200.edm.main as em
    _install_stub_kernels(monkeypatch.setattr)
    monkeypatch.setitem(_STUB_NOISE, 0, 4e-3)
    N, steps, b = 12, 6, 2
    g = torch.Generator().manual_seed(17)
    latents = torch.randn(b, 3, 8, 8, generator=g)
    pre = {}
    for i in range(steps):
        pre[f'pivot_{i}'] = torch.randn(b, 3, 8, 8, generator=g, dtype=torch.float64)
        pre[i] = torch.randn(b, 1, N, 3, 8, 8, generator=g, dtype=torch.float64)
    table = _StubTable(steps, noisy=set(range(steps)))
    scales = (torch.rand(steps, 1, N, generator=g) * 3).to(torch.float32)
    params = em.SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=_StubScorer())
    kw = dict(kappa=kappa) if kappa is not None else dict(delta=delta)
    x, rec = em.eps_greedy_search(_StubPreciseNet(), latents, None, params, table, precomputed_noise=pre, record=True,
                                  scale_table=scales, escalate=True, max_contenders=M, **kw)
    assert len(rec.indices) == steps
    flips = refined_rows = 0
    x_cur = latents.to(torch.float64) * table.t_steps[0]
    for i in range(steps):
        s16 = rec.scores[i]                                           # [N, b] as scored by the stub 16-bit engine
        cand = pre[f'pivot_{i}'].repeat(N, 1, 1, 1) + scales[i, 0].repeat_interleave(b).view(-1, 1, 1, 1) * \
            pre[i][:, 0].transpose(0, 1).reshape(N * b, 3, 8, 8) / pre[i][:, 0].transpose(0, 1).reshape(N * b, -1).norm(dim=1).view(-1, 1, 1, 1)
        xs, s_chk = _stub_step(x_cur, cand, i)
        assert torch.equal(s_chk.reshape(N, b), s16)
        want = _escalated_reference(s16, _true_score(xs).reshape(N, b), kappa, delta, M)
        assert rec.indices[i].tolist() == want, (i, rec.indices[i].tolist(), want)
        flips += sum(int(a != c) for a, c in zip(want, s16.argmax(dim=0).tolist()))
        if rec.refined[i] is not None:
            refined_rows += int((rec.refined[i] > float('-inf')).sum()) - int(((rec.refined[i] > float('-inf')).sum(0) == 1).sum())
        x_cur = rec.x_steps[i]
    assert sum(rec.escalated) == refined_rows and refined_rows > 0
    assert flips > 0, 'the stub perturbation is sized to flip near ties (3-5 of 12 image-rounds)'
    assert rec.mispredicted == 0

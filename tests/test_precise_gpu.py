"""The fp32-faithful ("precise", split-fp16) re-scoring engine on a B200: each kernel against a plain PyTorch fp64
reference of the same op, the whole network against the REFERENCE's fp32 output (tests/golden, written by
oracle/make_golden.py from /root/reference), and batch-position invariance.

Tolerance: the precise path exists to reproduce the reference's fp32 argmax, so its error must sit at the level of
fp32 summation-order noise: relative L2 of F_x < 4e-5 (measured 1.9e-5 on the full ADM-64; the bf16 engine: 7.4e-3, bound 2e-2)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import edm_oracle as O  # noqa: E402
from tests.helpers import load_golden, oracle_net  # noqa: E402

REL_TOL = 4e-5


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope='module')
def pkg():
    from diffusion_tts_b200 import build
    build.build()
    import diffusion_tts_b200.denoiser as den
    from diffusion_tts_b200 import ops, precise
    return den, ops, precise


def _split(t):
    """fp32/fp64 [..., C] -> split half [..., 2C] on the GPU."""
    t = t.float()
    hi = t.half()
    lo = (t - hi.float()).half()
    return torch.cat([hi, lo], dim=-1).contiguous().cuda()


def _join(t):
    C = t.shape[-1] // 2
    return t[..., :C].double() + t[..., C:].double()


@pytest.mark.parametrize('B,H,cin,cout,taps', [(2, 16, 64, 128, 9), (3, 8, 192, 192, 9), (1, 32, 128, 64, 1), (5, 8, 64, 64, 9)])
def test_gemm_prec_conv_matches_fp64(pkg, B, H, cin, cout, taps):
    den, ops, precise = pkg
    g = torch.Generator().manual_seed(B * 1000 + H)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, cin, H, H, generator=g) * 2
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(B, cout, H, H, generator=g)
    # the reference of the OP: inputs as the kernel sees them (the split representation of x and res is part of the input)
    xs, rs = _split(x.permute(0, 2, 3, 1)), _split(res.permute(0, 2, 3, 1))
    x64 = _join(xs.cpu()).permute(0, 3, 1, 2)
    r64 = _join(rs.cpu()).permute(0, 3, 1, 2)
    want = (torch.nn.functional.conv2d(x64, w.double(), bias.double(), padding=k // 2) + r64) * 0.75
    wp, acc_scale, segs = precise.pack_split([precise._conv_block(w)])
    flat = precise.flat_segs(segs)
    out = torch.empty(B, H, H, 2 * cout, dtype=torch.float16, device='cuda')
    P = ops.Plan()
    P.add_gemm_prec([xs], flat, wp.cuda(), cout, out, acc_scale=acc_scale, bias=bias.cuda(), residual=rs, out_scale=0.75)
    # the same GEMM with the layer's split-K policy (K slices on different SMs, partial tiles added in order)
    bn, splits = precise.split_k_policy(H * H, wp.shape[1], wp.shape[0] // 2)
    splits = max(splits, 3)
    out_sk = torch.empty_like(out)
    ws_ = torch.empty(splits * ((B * H * H + 127) // 128) * 128 * wp.shape[1], dtype=torch.float32, device='cuda')
    P.add_gemm_prec([xs], flat, wp.cuda(), cout, out_sk, acc_scale=acc_scale, bias=bias.cuda(), residual=rs, out_scale=0.75,
                    splits=splits, bn=bn, partial=ws_)
    P.run()
    torch.cuda.synchronize()
    assert _rel(_join(out_sk.cpu()).permute(0, 3, 1, 2), want) < 1.5e-5
    got = _join(out.cpu()).permute(0, 3, 1, 2)
    err = _rel(got, want)
    print(f'gemm_prec B={B} H={H} {cin}->{cout} taps={taps}: rel {err:.2e}')
    # the tensor core adds fp32 partial products with truncation, so the error grows with K (measured 6e-7 at K = 128,
    # 6e-6 at K = 1728): still ~3 orders of magnitude below bf16 storage (4e-3) -- see the whole-network figure below
    assert err < 1.5e-5


def test_gemm_prec_two_sources_and_fp32_out(pkg):
    """conv1 (3x3) + 1x1 skip over a concat of two tensors in one accumulator; and the 3-channel fp32 output conv."""
    den, ops, precise = pkg
    g = torch.Generator().manual_seed(5)
    B, H, c1, ca, cb, cout = 2, 16, 128, 64, 128, 128
    a1 = torch.randn(B, c1, H, H, generator=g)
    xa, xb = torch.randn(B, ca, H, H, generator=g), torch.randn(B, cb, H, H, generator=g)
    w1 = torch.randn(cout, c1, 3, 3, generator=g) / math.sqrt(9 * c1)
    ws = torch.randn(cout, ca + cb, 1, 1, generator=g) / math.sqrt(ca + cb)
    bias = torch.randn(cout, generator=g) * 0.1
    s1, sa, sb = (_split(t.permute(0, 2, 3, 1)) for t in (a1, xa, xb))
    j = lambda t: _join(t.cpu()).permute(0, 3, 1, 2)
    want = torch.nn.functional.conv2d(j(s1), w1.double(), bias.double(), padding=1) + \
        torch.nn.functional.conv2d(torch.cat([j(sa), j(sb)], 1), ws.double())
    wp, acc_scale, segs = precise.pack_split([precise._conv_block(w1), precise._conv_block(ws, 0, ca), precise._conv_block(ws, ca, ca + cb)])
    flat = precise.flat_segs(segs)
    out = torch.empty(B, H, H, 2 * cout, dtype=torch.float16, device='cuda')
    P = ops.Plan()
    P.add_gemm_prec([s1, sa, sb], flat, wp.cuda(), cout, out, acc_scale=acc_scale, bias=bias.cuda())
    # 3-channel fp32 output (Npad = 16)
    w3 = torch.randn(3, c1, 3, 3, generator=g) / math.sqrt(9 * c1)
    b3 = torch.randn(3, generator=g)
    wp3, sc3, segs3 = precise.pack_split([precise._conv_block(w3)], n_pad=16)
    out3 = torch.empty(B, H, H, 3, dtype=torch.float32, device='cuda')
    P.add_gemm_prec([s1], precise.flat_segs(segs3), wp3.cuda(), 3, out3, acc_scale=sc3, bias=b3.cuda())
    P.run()
    torch.cuda.synchronize()
    assert _rel(j(out), want) < 1.5e-5
    want3 = torch.nn.functional.conv2d(j(s1), w3.double(), b3.double(), padding=1)
    assert _rel(out3.cpu().permute(0, 3, 1, 2), want3) < 1.5e-5


@pytest.mark.parametrize('resample,two,film', [(0, False, True), (1, False, False), (2, True, False), (0, True, True)])
def test_gn_prec_matches_fp64(pkg, resample, two, film):
    den, ops, precise = pkg
    g = torch.Generator().manual_seed(7 + resample)
    B, H, C0, C1 = 3, 16, 128, (64 if two else 0)
    C = C0 + C1
    xs = [torch.randn(B, H, H, C0, generator=g) * 3 + 1] + ([torch.randn(B, H, H, C1, generator=g)] if two else [])
    sx = [_split(t) for t in xs]
    x64 = torch.cat([_join(t.cpu()) for t in sx], dim=-1).permute(0, 3, 1, 2)
    gamma, beta = torch.randn(C, generator=g) * 0.1 + 1, torch.randn(C, generator=g) * 0.1
    groups = min(32, C // 4)
    y = torch.nn.functional.group_norm(x64, groups, gamma.double(), beta.double(), eps=1e-5)
    fs = fh = None
    if film:
        fs, fh = torch.randn(1, C, generator=g) * 0.3, torch.randn(1, C, generator=g) * 0.3
        y = fh.double().view(1, C, 1, 1) + y * (fs.double().view(1, C, 1, 1) + 1)
    y = torch.nn.functional.silu(y)
    raw = x64
    if resample == 1:
        y, raw = (t.repeat_interleave(2, 2).repeat_interleave(2, 3) for t in (y, raw))
    elif resample == 2:
        y, raw = (torch.nn.functional.avg_pool2d(t, 2) for t in (y, raw))
    Ho = y.shape[2]
    out = torch.empty(B, Ho, Ho, 2 * C, dtype=torch.float16, device='cuda')
    raw_out = torch.empty_like(out)
    mr = torch.empty(B, groups, 2, device='cuda')
    P = ops.Plan()
    P.add_gn_prec(sx, groups, 1e-5, mr, gamma.cuda(), beta.cuda(), out, film_scale=fs.cuda() if film else None,
                  film_shift=fh.cuda() if film else None, b_emb=1, silu=True, resample=resample, raw_out=raw_out)
    P.run()
    torch.cuda.synchronize()
    assert _rel(_join(out.cpu()).permute(0, 3, 1, 2), y) < 2e-6
    assert _rel(_join(raw_out.cpu()).permute(0, 3, 1, 2), raw) < 1e-6


@pytest.mark.parametrize('L,heads', [(64, 2), (256, 3), (1024, 1)])
def test_attention_prec_matches_fp64(pkg, L, heads):
    den, ops, precise = pkg
    g = torch.Generator().manual_seed(L)
    B, C = 2, 64 * heads
    qkv = torch.randn(B * L, 3 * C, generator=g) * 1.5
    sq = _split(qkv)
    q64 = _join(sq.cpu()).view(B, L, 3, heads, 64)
    q, k, v = (q64[:, :, t].permute(0, 2, 1, 3) for t in range(3))                  # [B, heads, L, 64]
    w = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    want = (w @ v).permute(0, 2, 1, 3).reshape(B * L, C)
    out = torch.empty(B * L, 2 * C, dtype=torch.float16, device='cuda')
    P = ops.Plan()
    P.add_attention_prec(sq, out, B, heads, L, C)
    P.run()
    torch.cuda.synchronize()
    err = _rel(_join(out.cpu()), want)
    print(f'attention_prec L={L}: rel {err:.2e}')
    assert err < 2e-6


def _precise_F(den, g, case):
    net, spec, sd = oracle_net(g['cfg'], g['seed'])
    eng = den.B200Denoiser(sd, device='cuda')
    assert eng.supports_precise
    c_skip, c_out, c_in, c_noise = O.precond_coeffs(torch.tensor(case['sigma'], dtype=torch.float64))
    labels = case['labels'].cuda() if case['labels'] is not None else None
    b_emb = eng._distinct_rows(labels, case['x'].shape[0]) if labels is not None else 1
    F = eng.precise_engine.forward((c_in.float() * case['x']).cuda().contiguous(), c_noise.float().flatten().cuda(),
                                   labels[:b_emb] if labels is not None else None, b_emb=b_emb)
    return eng, F.cpu()


def test_precise_tiny_adm_matches_reference_golden(pkg):
    den, ops, precise = pkg
    g = load_golden('unet_tiny_adm.pt')
    for case in g['cases']:
        eng, F = _precise_F(den, g, case)
        err = _rel(F, case['F'])
        print('precise tiny ADM sigma', case['sigma'], 'rel err', err)
        assert err < REL_TOL


def test_precise_full_adm_matches_reference_golden(pkg):
    """ImageNet-64 ADM (295.9 M parameters) against the reference's own fp32 forward."""
    den, ops, precise = pkg
    g = load_golden('unet_full_adm.pt')
    case = g['cases'][0]
    eng, F = _precise_F(den, g, case)
    err = _rel(F, case['F'])
    # the bf16 engine on the same input, for the record
    sigma = torch.tensor(case['sigma'], dtype=torch.float64, device='cuda')
    D = eng(case['x'].cuda(), sigma, case['labels'].cuda()).cpu()
    c_skip, c_out, c_in, c_noise = O.precond_coeffs(torch.tensor(case['sigma'], dtype=torch.float64))
    print('full ADM rel err: precise %.2e, bf16 %.2e' % (err, _rel((D - c_skip * case['x']) / c_out, case['F'])))
    assert err < REL_TOL


def test_precise_batch_position_invariance(pkg):
    den, ops, precise = pkg
    g = load_golden('unet_tiny_adm.pt')
    case = g['cases'][1]
    net, spec, sd = oracle_net(g['cfg'], g['seed'])
    eng = den.B200Denoiser(sd, device='cuda')
    c_skip, c_out, c_in, c_noise = O.precond_coeffs(torch.tensor(case['sigma'], dtype=torch.float64))
    x1 = (c_in.float() * case['x'][:1]).cuda()
    lab = case['labels'][:1].cuda()
    F1 = eng.precise_engine.forward(x1, c_noise.float().flatten().cuda(), lab, b_emb=1).clone()
    F5 = eng.precise_engine.forward(x1.repeat(5, 1, 1, 1).contiguous(), c_noise.float().flatten().cuda(), lab, b_emb=1)
    for r in range(5):
        assert torch.equal(F5[r], F1[0])

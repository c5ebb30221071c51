"""Per-kernel parity on a B200: every CUDA kernel behind the C ABI against a plain PyTorch fp32
(U-Net leaves) or the fp64 oracle (sampler/scorer; bit-exact).  All `@pytest.mark.gpu`."""
import math

import pytest
import torch

from diffusion_tts_b200._lib import ACT_DTYPE as ACT  # noqa: E402  (the engine's 16-bit storage type)
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import edm_oracle as O  # noqa: E402


@pytest.fixture(scope='module')
def ops():
    from diffusion_tts_b200 import build, ops
    build.build()
    return ops


def _rel_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _nhwc(x):       # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(ACT)


def _pack_w(w):     # [Cout,Cin,k,k] -> bf16 [Cout, k*k*Cin] (tap-major, channel-minor)
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(ACT)


@pytest.mark.parametrize('B,H,Cin,Cout,k', [
    (2, 16, 64, 64, 1), (2, 16, 64, 128, 3), (3, 8, 128, 192, 3), (1, 8, 64, 64, 3), (2, 32, 192, 384, 3),
    (1, 64, 192, 192, 3), (2, 16, 576, 1728, 1), (2, 8, 768, 768, 3), (4, 8, 1536, 768, 3), (2, 64, 64, 256, 1),
])
def test_conv_gemm(ops, B, H, Cin, Cout, k):
    torch.manual_seed(0)
    dev = 'cuda'
    x = torch.randn(B, Cin, H, H, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, device=dev)
    xa, wp = _nhwc(x), _pack_w(w)
    out = torch.empty(B, H, H, Cout, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_gemm([xa], [(0, k * k, 0, Cin // 64)], wp, Cout, out, bias=bias)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(xa.float().permute(0, 3, 1, 2), wp.float().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2), bias,
                   padding=k // 2).permute(0, 2, 3, 1)
    err = _rel_err(out, ref)
    assert err < 6e-3, f'rel err {err}'


@pytest.mark.parametrize('B,H,C,N', [(3, 16, 128, 384), (5, 8, 192, 576), (2, 32, 64, 128), (64, 8, 768, 2304), (2, 32, 384, 1152),
                                     (1, 8, 64, 64), (4, 16, 320, 320)])
def test_gemm_with_groupnorm_in_the_operand_path_is_bit_identical(ops, B, H, C, N):
    """`qkv(norm2(x))` (networks.py:182-183): GroupNorm applied to the TMA-landed A stage in shared memory by the GEMM's
    transform warps == gn_apply (silu = 0) followed by the plain GEMM, bit for bit -- incl. two samples per 128-row tile
    (H*W = 64), a partial last tile (odd batch), column-sliced launches (N = 320 = 192 + 128) and >= 2 waves of tiles."""
    torch.manual_seed(B * 100 + H)
    dev = 'cuda'
    groups = 32
    x = (torch.randn(B, H, H, C, device=dev) * 3 + 0.5).to(ACT)
    w = (torch.randn(N, C, device=dev) / math.sqrt(C)).to(ACT)
    bias = torch.randn(N, device=dev)
    gamma, beta = torch.randn(C, device=dev) * 0.3 + 1, torch.randn(C, device=dev) * 0.2
    xs = x.float().view(B, H * H, groups, C // groups)
    mean = xs.mean(dim=(1, 3))
    rstd = (xs.var(dim=(1, 3), unbiased=False) + 1e-5).rsqrt()
    mr = torch.stack([mean, rstd], dim=-1).contiguous()                       # [B, groups, 2]
    a2 = torch.empty_like(x)
    out_ref, out = torch.empty(B, H, H, N, device=dev, dtype=ACT), torch.empty(B, H, H, N, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_gn_apply([x], groups, None, 1, 1e-5, gamma, beta, a2, silu=False, mean_rstd=mr)
    plan.add_gemm([a2], [(0, 1, 0, C // 64)], w, N, out_ref, bias=bias)
    plan.add_gemm([x], [(0, 1, 0, C // 64)], w, N, out, bias=bias, a_norm=(mr, gamma, beta, groups))
    out_rev = torch.empty_like(out)
    plan.add_gemm([x], [(0, 1, 0, C // 64)], w, N, out_rev, bias=bias, a_norm=(mr, gamma, beta, groups), reverse=True)
    plan.run()
    torch.cuda.synchronize()
    want = F.linear(((xs - mean[:, None, :, None]) * rstd[:, None, :, None]).view(B, H, H, C) * gamma + beta, w.float(), bias)
    assert _rel_err(out_ref, want) < 6e-3
    assert torch.equal(out, out_ref), float((out.float() - out_ref.float()).abs().max())
    assert torch.equal(out_rev, out_ref)


def test_conv_fused_skip_dual_source_residual(ops):
    """conv1(h) + skip1x1(cat[x0,x1]) in one accumulator, + bias, * skip_scale; and the
    no-skip-conv flavour with a residual."""
    torch.manual_seed(1)
    dev = 'cuda'
    B, H, C0, C1, Co = 2, 16, 128, 64, 128
    h = torch.randn(B, Co, H, H, device=dev)
    x0 = torch.randn(B, C0, H, H, device=dev)
    x1 = torch.randn(B, C1, H, H, device=dev)
    w1 = torch.randn(Co, Co, 3, 3, device=dev) / math.sqrt(Co * 9)
    ws = torch.randn(Co, C0 + C1, 1, 1, device=dev) / math.sqrt(C0 + C1)
    bias = torch.randn(Co, device=dev)
    ha, x0a, x1a = _nhwc(h), _nhwc(x0), _nhwc(x1)
    # segment-major K packing: [conv1 taps x Co | skip C0 | skip C1]
    wp = torch.cat([_pack_w(w1), ws[:, :C0, 0, 0].to(ACT), ws[:, C0:, 0, 0].to(ACT)], dim=1).contiguous()
    out = torch.empty(B, H, H, Co, device=dev, dtype=ACT)
    # sources: 0 = h ; the skip operands come from a 2nd launch-level source, so run as two plans:
    # (a) h + x0 as sources [conv1 | skip(x0)], (b) check dual-source concat conv separately.
    plan = ops.Plan()
    wpa = torch.cat([_pack_w(w1), ws[:, :C0, 0, 0].to(ACT)], dim=1).contiguous()
    plan.add_gemm([ha, x0a], [(0, 9, 0, Co // 64), (1, 1, 0, C0 // 64)], wpa, Co, out, bias=bias, out_scale=0.5)
    plan.run()
    ref = (F.conv2d(ha.float().permute(0, 3, 1, 2), w1.to(ACT).float(), None, padding=1) +
           F.conv2d(x0a.float().permute(0, 3, 1, 2), ws[:, :C0].to(ACT).float()) +
           bias.view(1, -1, 1, 1)) * 0.5
    assert _rel_err(out, ref.permute(0, 2, 3, 1)) < 6e-3
    # dual-source 3x3 conv over cat([x0, x1]) with a residual
    wc = torch.randn(Co, C0 + C1, 3, 3, device=dev) / math.sqrt((C0 + C1) * 9)
    wpc = torch.cat([_pack_w(wc[:, :C0]), _pack_w(wc[:, C0:])], dim=1).contiguous()
    res = _nhwc(torch.randn(B, Co, H, H, device=dev))
    out2 = torch.empty_like(out)
    plan2 = ops.Plan()
    plan2.add_gemm([x0a, x1a], [(0, 9, 0, C0 // 64), (1, 9, 0, C1 // 64)], wpc, Co, out2, bias=bias, residual=res)
    plan2.run()
    xcat = torch.cat([x0a, x1a], dim=3).float().permute(0, 3, 1, 2)
    ref2 = F.conv2d(xcat, wc.to(ACT).float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    assert _rel_err(out2, ref2) < 6e-3


def test_conv_small_n_fp32_out(ops):
    """out_conv: Cout=3 (weights padded to 16 rows), fp32 output."""
    torch.manual_seed(2)
    dev = 'cuda'
    B, H, Cin = 2, 16, 192
    x = torch.randn(B, Cin, H, H, device=dev)
    w = torch.randn(3, Cin, 3, 3, device=dev) / math.sqrt(Cin * 9)
    bias = torch.randn(3, device=dev)
    wp = torch.zeros(16, 9 * Cin, device=dev, dtype=ACT)
    wp[:3] = _pack_w(w)
    out = torch.empty(B, H, H, 3, device=dev, dtype=torch.float32)
    plan = ops.Plan()
    plan.add_gemm([_nhwc(x)], [(0, 9, 0, Cin // 64)], wp, 3, out, bias=bias)
    plan.run()
    ref = F.conv2d(_nhwc(x).float().permute(0, 3, 1, 2), w.to(ACT).float(), bias, padding=1).permute(0, 2, 3, 1)
    assert _rel_err(out, ref) < 2e-3


def test_first_conv_im2col(ops):
    torch.manual_seed(3)
    dev = 'cuda'
    B, H, Co = 2, 16, 64
    x = torch.randn(B, 3, H, H, device=dev)
    w = torch.randn(Co, 3, 3, 3, device=dev) / math.sqrt(27)
    bias = torch.randn(Co, device=dev)
    col = torch.empty(B, H, H, 64, device=dev, dtype=ACT)
    wp = torch.zeros(Co, 64, device=dev, dtype=ACT)
    wp[:, :27] = _pack_w(w)
    out = torch.empty(B, H, H, Co, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_im2col(x, col)
    plan.add_gemm([col], [(0, 1, 0, 1)], wp, Co, out, bias=bias)
    plan.run()
    ref = F.conv2d(x.to(ACT).float(), w.to(ACT).float(), bias, padding=1).permute(0, 2, 3, 1)
    assert _rel_err(out, ref) < 6e-3


@pytest.mark.parametrize('C0,C1,H,resample,film,silu', [
    (64, 0, 16, 0, False, True), (192, 0, 16, 0, True, True), (128, 64, 8, 0, False, True),
    (768, 576, 8, 0, True, True), (192, 0, 16, 2, False, True), (384, 0, 8, 1, False, True),
    (128, 0, 8, 0, False, False),
])
def test_groupnorm(ops, C0, C1, H, resample, film, silu):
    torch.manual_seed(4)
    dev = 'cuda'
    B, C = 3, C0 + C1
    groups = min(32, C // 4)
    x0 = _nhwc(torch.randn(B, C0, H, H, device=dev) * 2 + 0.5)
    xs = [x0]
    if C1:
        xs.append(_nhwc(torch.randn(B, C1, H, H, device=dev)))
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    b_emb = 1
    fs = torch.randn(b_emb, 2 * C, device=dev) * 0.3 if film else None
    splits = 4
    partial = torch.empty(B, splits, groups, 2, device=dev, dtype=torch.float64)
    Ho = H * 2 if resample == 1 else (H // 2 if resample == 2 else H)
    out = torch.empty(B, Ho, Ho, C, device=dev, dtype=ACT)
    raw = torch.empty_like(out)
    plan = ops.Plan()
    plan.add_gn_stats(xs, groups, partial, splits)
    plan.add_gn_apply(xs, groups, partial, splits, 1e-5, gamma, beta, out, film_scale=fs[:, :C] if film else None,
                      film_shift=fs[:, C:] if film else None, b_emb=b_emb, silu=silu, resample=resample, raw_out=raw)
    plan.run()
    xc = torch.cat(xs, dim=3).float().permute(0, 3, 1, 2)
    y = F.group_norm(xc, groups, gamma, beta, 1e-5)
    if film:
        y = torch.addcmul(fs[:, C:, None, None], y, fs[:, :C, None, None] + 1)
    if silu:
        y = F.silu(y)
    r = xc
    if resample == 1:
        y, r = O._resample(y, True, False), O._resample(r, True, False)
    elif resample == 2:
        y, r = O._resample(y, False, True), O._resample(r, False, True)
    assert _rel_err(out, y.permute(0, 2, 3, 1)) < 5e-3
    assert _rel_err(raw, r.permute(0, 2, 3, 1)) < 5e-3


@pytest.mark.parametrize('B,heads,L', [(2, 2, 64), (1, 6, 1024), (3, 9, 256), (2, 1, 128)])
def test_attention(ops, B, heads, L):
    torch.manual_seed(5)
    dev = 'cuda'
    C = heads * 64
    q = torch.randn(B, L, C, device=dev)
    k = torch.randn(B, L, C, device=dev)
    v = torch.randn(B, L, C, device=dev)
    qk = torch.cat([q, k], dim=2).to(ACT).contiguous()                  # [B, L, 2C]
    vt = v.to(ACT).reshape(B, L, heads, 64).permute(0, 2, 3, 1).contiguous()   # [B, heads, 64, L]
    out = torch.zeros(B, L, C, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_attention(qk.reshape(B * L, 2 * C), C, vt.reshape(B * heads * 64, L), out.reshape(B * L, C), B, heads, L)
    plan.run()
    qf = qk[..., :C].float().reshape(B, L, heads, 64).permute(0, 2, 1, 3)
    kf = qk[..., C:].float().reshape(B, L, heads, 64).permute(0, 2, 1, 3)
    vf = vt.float().permute(0, 1, 3, 2)                                             # [B, heads, L, 64]
    w = torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, dim=-1)
    ref = (w @ vf).permute(0, 2, 1, 3).reshape(B, L, C)
    assert _rel_err(out, ref) < 1e-2


@pytest.mark.parametrize('B,heads,L', [(2, 2, 64), (1, 6, 1024), (3, 9, 256), (2, 1, 128)])
def test_attention_row_major_v(ops, B, heads, L):
    """V consumed in place from the row-major qkv matrix (MN-major UMMA B operand): no V^T copy."""
    torch.manual_seed(15)
    dev = 'cuda'
    C = heads * 64
    qkv = torch.randn(B, L, 3 * C, device=dev).to(ACT).contiguous()
    out = torch.zeros(B, L, C, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_attention(qkv.reshape(B * L, 3 * C), C, None, out.reshape(B * L, C), B, heads, L, v_col0=2 * C)
    plan.run()
    qf, kf, vf = [qkv[..., i * C:(i + 1) * C].float().reshape(B, L, heads, 64).permute(0, 2, 1, 3) for i in range(3)]
    w = torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, dim=-1)
    ref = (w @ vf).permute(0, 2, 1, 3).reshape(B, L, C)
    assert _rel_err(out, ref) < 1e-2


@pytest.mark.parametrize('B,L', [(2, 256), (3, 64), (1, 128)])
def test_attention_head_dim_256(ops, B, L):
    """DDPM++ attention: one head of 256 channels (networks.py:263), exact two-pass softmax in TMEM."""
    torch.manual_seed(16)
    dev = 'cuda'
    C = 256
    qkv = torch.randn(B, L, 3 * C, device=dev).to(ACT).contiguous()
    out = torch.zeros(B, L, C, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_attention(qkv.reshape(B * L, 3 * C), C, None, out.reshape(B * L, C), B, 1, L, v_col0=2 * C, head_dim=256)
    plan.run()
    q, k, v = [qkv[..., i * C:(i + 1) * C].float() for i in range(3)]
    w = torch.softmax(q @ k.transpose(-1, -2) / 16.0, dim=-1)
    assert _rel_err(out, w @ v) < 1e-2


@pytest.mark.parametrize('B,H,C0,C1,Cout,k,pre', [
    (2, 16, 128, 0, 192, 3, False), (3, 8, 64, 64, 128, 1, False), (2, 32, 64, 0, 384, 3, True), (5, 8, 192, 128, 768, 3, False),
])
def test_gemm_epilogue_gn_stats_and_finalize(ops, B, H, C0, C1, Cout, k, pre):
    """The GEMM epilogue leaves per-channel (sum, sumsq) of the STORED bf16 output per 64-row half tile; gn_finalize
    turns the statistics of up to two concatenated tensors into (mean, rstd); gn_apply consumes them.  Checked against
    sums recomputed from the stored tensor (tight) and against torch group_norm of the concat (GN tolerance)."""
    torch.manual_seed(6)
    dev = 'cuda'
    outs, stats = [], []
    for Cin in [c for c in (C0, C1) if c]:
        x = _nhwc(torch.randn(B, Cin, H, H, device=dev))
        w = torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)
        bias = torch.randn(Cout, device=dev)
        res = _nhwc(torch.randn(B, Cout, H, H, device=dev))
        out = torch.empty(B, H, H, Cout, device=dev, dtype=ACT)
        st = torch.full((B * H * H // 64, Cout, 2), float('nan'), device=dev)
        plan = ops.Plan()
        plan.add_gemm([x], [(0, k * k, 0, Cin // 64)], _pack_w(w), Cout, out, bias=bias, residual=res, out_scale=0.7,
                      gn_stats=st)
        plan.run()
        torch.cuda.synchronize()
        ref = (F.conv2d(x.float().permute(0, 3, 1, 2), _pack_w(w).float().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2), bias,
                        padding=k // 2).permute(0, 2, 3, 1) + res.float()) * 0.7
        assert _rel_err(out, ref) < 6e-3
        o64 = out.double().reshape(-1, 64, Cout)
        assert torch.allclose(st[..., 0].double(), o64.sum(1), rtol=1e-5, atol=1e-4)
        assert torch.allclose(st[..., 1].double(), (o64 * o64).sum(1), rtol=1e-5, atol=1e-4)
        outs.append(out)
        stats.append(st)
    C = Cout * len(outs)
    groups = min(32, C // 4)
    pa = torch.randn(1, C, device=dev) if pre else None
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    mr = torch.empty(B, groups, 2, device=dev)
    y = torch.empty(B, H, H, C, device=dev, dtype=ACT)
    plan = ops.Plan()
    plan.add_gn_finalize(stats, [Cout] * len(outs), B, H * H, groups, 1e-5, mr, pre_add=pa)
    plan.add_gn_apply(outs, groups, None, 1, 1e-5, gamma, beta, y, pre_add=pa, silu=True, mean_rstd=mr)
    plan.run()
    xc = torch.cat(outs, dim=3).float().permute(0, 3, 1, 2)
    if pre:
        xc = xc + pa.view(1, C, 1, 1)
    xg = xc.double().reshape(B, groups, -1)
    assert torch.allclose(mr[..., 0].double(), xg.mean(2), rtol=1e-4, atol=1e-5)
    assert torch.allclose(mr[..., 1].double(), 1.0 / torch.sqrt(xg.var(2, unbiased=False) + 1e-5), rtol=1e-4)
    ref = F.silu(F.group_norm(xc, groups, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert _rel_err(y, ref) < 5e-3


@pytest.mark.parametrize('B,H,C0,C1,resample,film,silu,pre,reverse', [
    (3, 8, 768, 0, 0, True, True, False, False), (2, 8, 768, 768, 0, False, True, False, True),
    (5, 16, 576, 0, 0, False, False, False, False), (3, 16, 576, 576, 0, True, True, False, True),
    (2, 16, 576, 384, 1, False, True, False, False), (3, 16, 384, 0, 2, True, True, False, False),
    (2, 8, 64, 0, 0, False, True, True, False), (4, 16, 192, 64, 0, True, True, True, False),
    (64, 8, 768, 0, 0, True, True, False, True),
])
def test_groupnorm_cluster_launch_is_bit_identical(ops, B, H, C0, C1, resample, film, silu, pre, reverse):
    """gn_norm_cluster_kernel (finalize + apply in one 8-CTA-cluster launch per sample, H*W <= 256) against the two
    separate kernels on the same epilogue statistics: every output bit and (mean, rstd) must agree (concat sources,
    FiLM, pre_add, both resample modes, partial last warps: C = 576 -> 216 threads, reversed walk)."""
    torch.manual_seed(12)
    dev = 'cuda'
    chans = [c for c in (C0, C1) if c]
    C = sum(chans)
    groups = min(32, C // 4)
    xs = [(torch.randn(B, H, H, c, device=dev) * 1.5 + 0.3).to(ACT) for c in chans]
    stats = []
    for x in xs:
        x64 = x.float().reshape(-1, 64, x.shape[3])
        stats.append(torch.stack([x64.sum(1), (x64 * x64).sum(1)], dim=2).contiguous())
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    pa = torch.randn(1, C, device=dev) if pre else None
    fs = torch.randn(1, 2 * C, device=dev) * 0.3 if film else None
    Ho = H * 2 if resample == 1 else (H // 2 if resample == 2 else H)
    res = {}
    for fused in (False, True):
        mr = torch.full((B, groups, 2), float('nan'), device=dev)
        out = torch.full((B, Ho, Ho, C), float('nan'), device=dev, dtype=ACT)
        raw = torch.full((B, Ho, Ho, C), float('nan'), device=dev, dtype=ACT)
        kw = dict(pre_add=pa, film_scale=fs[:, :C] if film else None, film_shift=fs[:, C:] if film else None, b_emb=1,
                  silu=silu, resample=resample, raw_out=raw, reverse=reverse)
        plan = ops.Plan()
        old = ops.GN_CLUSTER
        ops.GN_CLUSTER = fused
        try:
            plan.add_gn_norm(stats, xs, groups, 1e-5, mr, gamma, beta, out, **kw)
        finally:
            ops.GN_CLUSTER = old
        assert plan.kinds == (['gn_norm'] if fused else ['gn_finalize', 'gn_apply'])
        plan.run()
        torch.cuda.synchronize()
        res[fused] = (mr, out, raw)
    for a, b in zip(res[False], res[True]):
        assert not torch.isnan(a.float()).any()
        assert torch.equal(a, b)
    xc = torch.cat(xs, dim=3).float().permute(0, 3, 1, 2)
    if pre:
        xc = xc + pa.view(1, C, 1, 1)
    y = F.group_norm(xc, groups, gamma, beta, 1e-5)
    if film:
        y = torch.addcmul(fs[:, C:, None, None], y, fs[:, :C, None, None] + 1)
    if silu:
        y = F.silu(y)
    if resample:
        y = O._resample(y, resample == 1, resample == 2)
    assert _rel_err(res[True][1], y.permute(0, 2, 3, 1)) < 5e-3


def test_linear(ops):
    torch.manual_seed(7)
    dev = 'cuda'
    x = torch.randn(3, 200, device=dev)
    w = torch.randn(77, 200, device=dev)
    b = torch.randn(77, device=dev)
    add = torch.randn(3, 77, device=dev)
    out = torch.empty(3, 77, device=dev)
    plan = ops.Plan()
    plan.add_linear(x, w, out, bias=b, add=add, act=1)
    plan.run()
    ref = F.silu(x @ w.t() + b + add)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ sampler / scorer: bit-exact vs oracle
def test_sampler_kernels_bit_exact(ops):
    torch.manual_seed(8)
    dev = 'cuda'
    b, N, C, H = 2, 5, 3, 16
    R = N * b
    t_steps = O.karras_schedule(18)
    for i in (3, 17):
        t_cur, t_next = t_steps[i], t_steps[i + 1]
        gamma = O.churn_gamma(t_cur, 18, 40, 0.05, 50)
        t_hat = t_cur + gamma * t_cur
        x_cur = torch.randn(b, C, H, H, dtype=torch.float64) * t_cur
        eps = torch.randn(R, C, H, H, dtype=torch.float64)
        F1 = torch.randn(R, C, H, H)
        F2 = torch.randn(R, C, H, H)

        class FakeNet:      # returns the canned "network outputs" through the real preconditioning
            def __init__(self):
                self.calls = 0

            def __call__(self, x, sigma, labels):
                Fx = F1 if self.calls == 0 else F2
                self.calls += 1
                c_skip, c_out, c_in, _ = O.precond_coeffs(sigma)
                self.last_in = c_in * x.to(torch.float32)
                return c_skip * x.to(torch.float32) + c_out * Fx

        net = FakeNet()
        x_next_ref, den_ref = O.heun_step(net, x_cur.repeat(N, 1, 1, 1), t_cur, t_next, i, eps, None, num_steps=18,
                                          S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
        u8_ref = O.quantize_u8(den_ref)
        score_ref = O.brightness_score(u8_ref)

        s = ((t_hat ** 2 - t_cur ** 2).sqrt() * 1.003).item()
        cs1, co1, ci1, _ = [c.item() for c in O.precond_coeffs(t_hat)]
        x_hat, net_in = ops.heun_pre(x_cur.to(dev), eps.to(dev), s, ci1)
        x_hat_ref = x_cur.repeat(N, 1, 1, 1) + (t_hat ** 2 - t_cur ** 2).sqrt() * 1.003 * eps
        assert torch.equal(x_hat.cpu(), x_hat_ref)
        assert torch.equal(net_in.cpu(), torch.tensor(ci1) * x_hat_ref.to(torch.float32))
        F1d = F1.permute(0, 2, 3, 1).contiguous().to(dev)
        F2d = F2.permute(0, 2, 3, 1).contiguous().to(dev)
        dt = (t_next - t_hat).item()
        if i < 17:
            cs2, co2, ci2, _ = [c.item() for c in O.precond_coeffs(t_next)]
            net_in2 = ops.heun_mid(x_hat, F1d, cs1, co1, t_hat.item(), dt, ci2)
            assert torch.equal(net_in2.cpu(), net.last_in)
            x_next, u8, sums = ops.heun_post(x_hat, F1d, F2d, cs1, co1, t_hat.item(), dt, cs2, co2, t_next.item(),
                                             want_u8=True)
        else:
            x_next, u8, sums = ops.heun_post(x_hat, F1d, None, cs1, co1, t_hat.item(), dt, want_u8=True)
        assert torch.equal(x_next.cpu(), x_next_ref)
        assert torch.equal(u8.cpu(), u8_ref)
        scores = ops.brightness_from_sums(sums, C, H * H)
        assert (scores.cpu() - score_ref).abs().max() <= 1.2e-7
        idx = ops.argmax_first(scores.reshape(N, b))
        assert torch.equal(idx.cpu(), O.argmax_first(score_ref.reshape(N, b)))


def test_heun_pre_with_fp32_noise_follows_torch_promotion(ops):
    """The MCTS depth noises are fp32 (edm/main.py:445): torch evaluates `sqrt(...) * S_noise * eps_i` (0-dim fp64 x fp32
    tensor) as an fp32 product and adds it to the fp64 state.  Bit-exact against that expression."""
    g = torch.Generator().manual_seed(9)
    b, R, C, H = 2, 6, 3, 8
    x_cur = torch.randn(b, C, H, H, generator=g, dtype=torch.float64) * 40
    eps32 = torch.randn(R, C, H, H, generator=g)
    t_hat, t_cur = torch.tensor(61.7, dtype=torch.float64), torch.tensor(57.3, dtype=torch.float64)
    scale = (t_hat ** 2 - t_cur ** 2).sqrt() * 1.003
    term = scale * eps32
    assert term.dtype == torch.float32
    want = x_cur.repeat(R // b, 1, 1, 1) + term
    assert want.dtype == torch.float64
    c_in = torch.tensor(0.0161, dtype=torch.float32)
    x_hat, net_in = ops.heun_pre(x_cur.cuda(), eps32.cuda(), float(scale), float(c_in))
    assert torch.equal(x_hat.cpu(), want)
    assert torch.equal(net_in.cpu(), c_in * want.to(torch.float32))
    x_hat64, _ = ops.heun_pre(x_cur.cuda(), eps32.double().cuda(), float(scale), float(c_in))
    assert not torch.equal(x_hat64.cpu(), want)                    # the fp64 product differs in the last bits


def test_scorer_argmax_ties_and_keys(ops):
    dev = 'cuda'
    g = torch.Generator().manual_seed(9)
    imgs = torch.randint(0, 256, (40, 3, 64, 64), generator=g, dtype=torch.uint8)
    imgs[7] = imgs[3]
    sums = ops.channel_sums_u8(imgs.to(dev))
    scores = ops.brightness_from_sums(sums, 3, 64 * 64).cpu()
    ref = O.brightness_score(imgs)
    assert (scores - ref).abs().max() <= 1.2e-7 and scores[7] == scores[3]
    # 4-channel fallback (sd/scorers.py:66-67)
    im4 = torch.randint(0, 256, (6, 4, 64, 64), generator=g, dtype=torch.uint8)
    s4 = ops.brightness_from_sums(ops.channel_sums_u8(im4.to(dev)), 4, 64 * 64).cpu()
    assert (s4 - O.brightness_score(im4)).abs().max() <= 1.2e-7
    # exact ties -> first index; keys reduce (max) to the same winner across shards
    sc = torch.tensor([[1., 0.], [3., 0.], [3., 0.], [2., 0.], [3., 0.], [-1., 0.]], device=dev)
    idx, key = ops.argmax_first(sc, want_key=True)
    assert idx.tolist() == [1, 0]
    idx_a, key_a = ops.argmax_first(sc[:3].contiguous(), idx_base=0, want_key=True)
    idx_b, key_b = ops.argmax_first(sc[3:].contiguous(), idx_base=3, want_key=True)
    merged = torch.maximum(key_a, key_b)          # what ncclAllReduce(max) computes (keys are < 2^63)
    assert torch.equal(merged, key)
    assert ((0xFFFFFFFF - (merged & 0xFFFFFFFF))).tolist() == [1, 0]


def test_candidates_and_gather(ops):
    torch.manual_seed(10)
    dev = 'cuda'
    b, N, C, H = 2, 6, 3, 16
    R = N * b
    pivot = torch.randn(b, C, H, H, dtype=torch.float64)
    dirs = torch.randn(R, C, H, H, dtype=torch.float64)
    fresh = torch.randn(R, C, H, H, dtype=torch.float64)
    lam = 0.15 * math.sqrt(3 * 64 * 64)
    scales = [O.candidate_scale_fp32((r * 37 % 1000) / 1000.0, lam) for r in range(R)]
    mask = torch.zeros(R, dtype=torch.uint8)
    mask[[2, 3, 9]] = 1
    # oracle (candidate n covers rows n*b .. n*b+b-1)
    ref = []
    for n in range(N):
        rows = slice(n * b, (n + 1) * b)
        if mask[n * b]:
            ref.append(fresh[rows])
        else:
            ref.append(O.make_candidates(pivot, [dirs[rows]], [scales[n * b]], [None]))
    mask = mask.reshape(N, b)
    mask[:] = mask[:, :1]
    mask = mask.reshape(R)
    ref = torch.cat(ref)
    # strict mode: norms from the same torch call the reference makes
    norms = torch.norm(dirs, p=2, dim=(1, 2, 3))
    sc = torch.stack([scales[(r // b) * b] for r in range(R)]).to(torch.float32)
    cand = ops.make_candidates(pivot.to(dev), dirs.to(dev), norms.to(dev), sc.to(dev), mask.to(dev), fresh.to(dev))
    assert torch.equal(cand.cpu(), ref)
    # fast mode: in-kernel norms agree to 1 ulp-ish
    fast = ops.direction_norms(dirs.to(dev)).cpu()
    assert torch.allclose(fast, norms, rtol=1e-14, atol=0)
    idx = torch.tensor([4, 1], device=dev)
    got = ops.gather_rows(cand.reshape(N, b, C, H, H), idx).cpu()
    assert torch.equal(got, torch.stack([ref.reshape(N, b, C, H, H)[4, 0], ref.reshape(N, b, C, H, H)[1, 1]]))
    assert torch.equal(ops.quantize_u8(cand).cpu(), O.quantize_u8(ref))


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 8, 8, 64, 128), (3, 16, 16, 128, 192), (1, 32, 64, 64, 320), (1, 4, 256, 64, 64)])
def test_conv_fused_nearest_upsample(ops, B, H, W, Cin, Cout):
    """out = conv3x3(nearest_upsample_2x(x)) from the LOW-res x as four 2x2-tap phase launches (pre-summed weights): against
    torch on the same bf16-rounded phase weights, plus the high-res GroupNorm statistics side band."""
    torch.manual_seed(3)
    x = torch.randn(B, H, W, Cin, device='cuda').to(ACT)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda') / math.sqrt(Cin * 9)
    bias = torch.randn(Cout, device='cuda')
    wp = ops.pack_conv_up2(w).cuda()
    out = torch.zeros(B, 2 * H, 2 * W, Cout, device='cuda', dtype=ACT)
    st = torch.zeros(B * 4 * H * W // 64, Cout, 2, device='cuda')
    plan = ops.Plan()
    plan.add_gemm([x], [(0, 9, 0, Cin // 64)], wp, Cout, out, bias=bias, gn_stats=st, upsample2x=True)
    plan.run()
    torch.cuda.synchronize()
    # exact-arithmetic reference of the same decomposition: phase kernels (bf16-rounded) applied to the padded low-res input
    xf = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    ref = torch.zeros(B, Cout, 2 * H, 2 * W, device='cuda')
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        k = wp[ph].float().reshape(Cout, 4, Cin)
        for a in range(2):
            for c in range(2):
                ref[:, :, py::2, px::2] += torch.einsum('oc,bchw->bohw', k[:, a * 2 + c], xf[:, :, py + a:py + a + H, px + c:px + c + W])
    ref = (ref + bias.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    assert _rel_err(out, ref) < 6e-3
    # and against the un-fused definition (differs only by the bf16 rounding of the pre-summed weights)
    full = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode='nearest'), w, bias, padding=1).permute(0, 2, 3, 1)
    assert _rel_err(out, full) < 1e-2
    per_img = out.float().reshape(B, -1, Cout)
    s_img = st.reshape(B, -1, Cout, 2)
    assert torch.allclose(s_img[..., 0].sum(1), per_img.sum(1), rtol=1e-3, atol=5e-2)
    assert torch.allclose(s_img[..., 1].sum(1), (per_img ** 2).sum(1), rtol=1e-3, atol=5e-2)

"""Thin Python wrappers over the C ABI: sampler/scorer ops on torch CUDA tensors and the
`Plan` builder for the U-Net engine.  Everything here launches hand-written sm_100a kernels
from libb200ns.so on torch's current stream; nothing falls back to PyTorch math."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import ACT_DTYPE
from . import torch_ops as T

# The sampler / scorer entry points and the plan runner go through their torch.library registrations
# (torch.ops.b200ns.*, torch_ops.py); B200NS_TORCH_OPS=0 calls the C ABI through ctypes directly.
USE_TORCH_OPS = os.environ.get('B200NS_TORCH_OPS', '1') != '0'
# 1 = gn_finalize + gn_apply as one thread-block-cluster launch at H*W <= 256 (bit-identical).  Measured on B200 at batch 64
# (profiles/r02_gn_cluster_ab.txt): 4.7 us saved per 8x8 norm, nothing at 16x16, NFE graph 16.12-16.34 ms against
# 16.05-16.19 ms with the two launches -- off by default.
GN_CLUSTER = os.environ.get('B200NS_GN_CLUSTER', '0') == '1'

# number of libb200ns kernel launches issued through this module (bench.py reports it)
LAUNCHES = [0]


def _count(n: int = 1):
    LAUNCHES[0] += n


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('b200 ops need CUDA tensors (no CPU fallback)')


def _c(t: torch.Tensor, dtype) -> torch.Tensor:
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f'expected contiguous {dtype} tensor, got {t.dtype} contiguous={t.is_contiguous()}')
    return t


# ------------------------------------------------------------------ sampler / scorer
def heun_pre(x_cur: torch.Tensor, eps: torch.Tensor, s: float, c_in: float, *, x_hat=None, net_in=None):
    """x_hat = x_cur + s*eps; net_in = c_in*fp32(x_hat).  x_cur [b,C,H,W] fp64, eps [R,C,H,W] fp64 -- or fp32, in which
    case the noise term is the fp32 product fp32(s)*eps like torch's own type promotion (edm/main.py:85 with the fp32 MCTS
    depth noises of :445).  `x_hat` / `net_in` may be preallocated (e.g. the U-Net engine's static input buffer)."""
    _chk_cuda(x_cur, eps)
    _c(x_cur, torch.float64), _c(eps, torch.float32 if eps.dtype == torch.float32 else torch.float64)
    R, b = eps.shape[0], x_cur.shape[0]
    E = eps[0].numel()
    x_hat = torch.empty(eps.shape, dtype=torch.float64, device=eps.device) if x_hat is None else _c(x_hat, torch.float64)
    net_in = torch.empty(eps.shape, dtype=torch.float32, device=eps.device) if net_in is None else _c(net_in, torch.float32)
    if USE_TORCH_OPS:
        T.OPS.heun_pre_(x_cur, eps, x_hat, net_in, float(s), float(c_in))
    else:
        fn = L.lib().b200ns_heun_pre_f32noise if eps.dtype == torch.float32 else L.lib().b200ns_heun_pre
        L.check(fn(L.ptr(x_cur), L.ptr(eps), L.ptr(x_hat), L.ptr(net_in), R, b, E, float(s), float(c_in), L.cur_stream()),
                'heun_pre')
    _count()
    return x_hat, net_in


def heun_mid(x_hat: torch.Tensor, F1: torch.Tensor, c_skip, c_out, t_hat, dt, c_in_next, want_x_eul=False, *,
             net_in2=None):
    """Euler half step.  F1 fp32 NHWC [R,H,W,C] (U-Net output)."""
    _chk_cuda(x_hat, F1)
    _c(x_hat, torch.float64), _c(F1, torch.float32)
    R, Cc, H, W = x_hat.shape
    net_in2 = (torch.empty(x_hat.shape, dtype=torch.float32, device=x_hat.device) if net_in2 is None
               else _c(net_in2, torch.float32))
    x_eul = torch.empty_like(x_hat) if want_x_eul else None
    if USE_TORCH_OPS:
        T.OPS.heun_mid_(x_hat, F1, net_in2, x_eul, float(c_skip), float(c_out), float(t_hat), float(dt), float(c_in_next))
    else:
        L.check(L.lib().b200ns_heun_mid(L.ptr(x_hat), L.ptr(F1), L.ptr(net_in2), L.ptr(x_eul), R, Cc, H * W, float(c_skip),
                                        float(c_out), float(t_hat), float(dt), float(c_in_next), L.cur_stream()), 'heun_mid')
    _count()
    return (net_in2, x_eul) if want_x_eul else net_in2


def heun_post(x_hat, F1, F2, c_skip1, c_out1, t_hat, dt, c_skip2=0.0, c_out2=0.0, t_next=1.0,
              want_x_next=True, want_u8=False, want_sums=True):
    """Heun correction (F2 None = last step) + x0 -> uint8 + integer channel sums."""
    _chk_cuda(x_hat, F1, F2)
    _c(x_hat, torch.float64), _c(F1, torch.float32)
    if F2 is not None:
        _c(F2, torch.float32)
    R, Cc, H, W = x_hat.shape
    dev = x_hat.device
    x_next = torch.empty_like(x_hat) if want_x_next else None
    u8 = torch.empty(x_hat.shape, dtype=torch.uint8, device=dev) if want_u8 else None
    sums = torch.empty((R, 4), dtype=torch.int32, device=dev) if want_sums else None
    if USE_TORCH_OPS:
        T.OPS.heun_post_(x_hat, F1, F2, x_next, u8, sums, float(c_skip1), float(c_out1), float(t_hat), float(dt),
                         float(c_skip2), float(c_out2), float(t_next))
    else:
        L.check(L.lib().b200ns_heun_post(L.ptr(x_hat), L.ptr(F1), L.ptr(F2), L.ptr(x_next), L.ptr(u8), L.ptr(sums), R, Cc,
                                         H * W, float(c_skip1), float(c_out1), float(t_hat), float(dt), float(c_skip2),
                                         float(c_out2), float(t_next), L.cur_stream()), 'heun_post')
    _count()
    return x_next, u8, sums


def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    _chk_cuda(x)
    _c(x, torch.float64)
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if USE_TORCH_OPS:
        T.OPS.quantize_u8_(x, out)
    else:
        L.check(L.lib().b200ns_quantize_u8(L.ptr(x), L.ptr(out), x.numel(), L.cur_stream()), 'quantize_u8')
    _count()
    return out


def channel_sums_u8(img: torch.Tensor) -> torch.Tensor:
    _chk_cuda(img)
    _c(img, torch.uint8)
    M, Cc = img.shape[0], img.shape[1]
    HW = img[0, 0].numel()
    sums = torch.empty((M, 4), dtype=torch.int32, device=img.device)
    if USE_TORCH_OPS:
        T.OPS.channel_sums_u8_(img, sums)
    else:
        L.check(L.lib().b200ns_channel_sums_u8(L.ptr(img), L.ptr(sums), M, Cc, HW, L.cur_stream()), 'channel_sums_u8')
    _count()
    return sums


def brightness_from_sums(sums: torch.Tensor, Cc: int, HW: int) -> torch.Tensor:
    _chk_cuda(sums)
    M = sums.shape[0]
    scores = torch.empty((M,), dtype=torch.float32, device=sums.device)
    if USE_TORCH_OPS:
        T.OPS.brightness_from_sums_(sums, scores, Cc, HW)
    else:
        L.check(L.lib().b200ns_brightness_from_sums(L.ptr(sums), L.ptr(scores), M, Cc, HW, L.cur_stream()), 'brightness')
    _count()
    return scores


def argmax_first(scores: torch.Tensor, idx_base: int = 0, want_key: bool = False):
    """scores [N,b] fp32 -> idx [b] int64 (first maximal n) and optionally the packed u64 key."""
    _chk_cuda(scores)
    _c(scores, torch.float32)
    N, b = scores.shape
    idx = torch.empty((b,), dtype=torch.int64, device=scores.device)
    key = torch.empty((b,), dtype=torch.int64, device=scores.device) if want_key else None
    if USE_TORCH_OPS:
        T.OPS.argmax_first_(scores, idx_base, idx, key)
    else:
        L.check(L.lib().b200ns_argmax_first(L.ptr(scores), N, b, idx_base, L.ptr(idx), L.ptr(key), L.cur_stream()), 'argmax')
    _count()
    return (idx, key) if want_key else idx


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """src [N,b,...] fp64, idx [b] -> [b,...]."""
    _chk_cuda(src, idx)
    _c(src, torch.float64), _c(idx, torch.int64)
    N, b = src.shape[0], src.shape[1]
    E = src[0, 0].numel()
    dst = torch.empty(src.shape[1:], dtype=torch.float64, device=src.device)
    if USE_TORCH_OPS:
        T.OPS.gather_rows_(src, idx, dst)
    else:
        L.check(L.lib().b200ns_gather_rows(L.ptr(src), L.ptr(idx), L.ptr(dst), N, b, E, L.cur_stream()), 'gather_rows')
    _count()
    return dst


def direction_norms(dirs: torch.Tensor) -> torch.Tensor:
    _chk_cuda(dirs)
    _c(dirs, torch.float64)
    R = dirs.shape[0]
    norms = torch.empty((R,), dtype=torch.float64, device=dirs.device)
    if USE_TORCH_OPS:
        T.OPS.direction_norms_(dirs, norms)
    else:
        L.check(L.lib().b200ns_direction_norms(L.ptr(dirs), L.ptr(norms), R, dirs[0].numel(), L.cur_stream()), 'norms')
    _count()
    return norms


def make_candidates(pivot, dirs, norms, scale, fresh_mask=None, fresh=None) -> torch.Tensor:
    """pivot [b,...] fp64; dirs [R,...] fp64; norms [R] fp64; scale [R] fp32; fresh_mask [R] u8; fresh [R,...]."""
    _chk_cuda(pivot, dirs, norms, scale, fresh_mask, fresh)
    _c(pivot, torch.float64), _c(dirs, torch.float64), _c(norms, torch.float64), _c(scale, torch.float32)
    R, b = dirs.shape[0], pivot.shape[0]
    E = dirs[0].numel()
    cand = torch.empty_like(dirs)
    if USE_TORCH_OPS:
        T.OPS.make_candidates_(pivot, dirs, norms, scale, fresh_mask, fresh, cand)
    else:
        L.check(L.lib().b200ns_make_candidates(L.ptr(pivot), L.ptr(dirs), L.ptr(norms), L.ptr(scale), L.ptr(fresh_mask),
                                               L.ptr(fresh), L.ptr(cand), R, b, E, L.cur_stream()), 'make_candidates')
    _count()
    return cand


def ddim_cfg_step(eps_pair: torch.Tensor, sample: torch.Tensor, noise: Optional[torch.Tensor], per_parent: int, guidance: float,
                  sqrt_beta_t: float, sqrt_alpha_t: float, sqrt_alpha_prev: float, dir_coef: float, std_dev: float, *,
                  net_in: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CFG + DDIM step (variance noise supplied).  eps_pair fp32 NHWC [2P,H,W,C] = UNet output for [uncond; cond];
    sample fp32 NCHW [P,C,H,W]; noise fp32 NCHW [P*per_parent,C,H,W] or None.  Returns prev [R,C,H,W]; fills net_in
    ([2R,C,H,W], both CFG halves) when given."""
    _chk_cuda(eps_pair, sample, noise, net_in)
    _c(eps_pair, torch.float32), _c(sample, torch.float32)
    P, Cc, H, W = sample.shape
    R = P * per_parent
    if eps_pair.shape[0] != 2 * P:
        raise RuntimeError('ddim_cfg_step: eps_pair must hold the uncond and cond halves of the P parents')
    if noise is not None:
        _c(noise, torch.float32)
    prev = torch.empty((R, Cc, H, W), dtype=torch.float32, device=sample.device)
    eu, et = eps_pair[:P], eps_pair[P:]
    L.check(L.lib().b200ns_ddim_cfg_step(L.ptr(eu), L.ptr(et), L.ptr(sample), L.ptr(noise), L.ptr(prev),
                                         L.ptr(_c(net_in, torch.float32)) if net_in is not None else None, R, per_parent, Cc,
                                         H * W, float(guidance), float(sqrt_beta_t), float(sqrt_alpha_t),
                                         float(sqrt_alpha_prev), float(dir_coef), float(std_dev), L.cur_stream()), 'ddim_cfg_step')
    _count()
    return prev


def ddim_x0_score(eps_pair: torch.Tensor, cand: torch.Tensor, guidance: float, sqrt_beta_t: float, sqrt_alpha_t: float, *,
                  want_x0: bool = False):
    """Guided eps of the 2nd UNet call -> pred_x0 -> uint8 -> mean(u8/255).  eps_pair fp32 NHWC [2R,H,W,C]; cand fp32
    NCHW [R,C,H,W].  Returns (scores fp32 [R], integer sums int32 [R], pred_x0 or None)."""
    _chk_cuda(eps_pair, cand)
    _c(eps_pair, torch.float32), _c(cand, torch.float32)
    R, Cc, H, W = cand.shape
    scores = torch.empty((R,), dtype=torch.float32, device=cand.device)
    sums = torch.empty((R,), dtype=torch.int32, device=cand.device)
    x0 = torch.empty_like(cand) if want_x0 else None
    L.check(L.lib().b200ns_ddim_x0_score(L.ptr(eps_pair[:R]), L.ptr(eps_pair[R:]), L.ptr(cand), L.ptr(x0), L.ptr(sums),
                                         L.ptr(scores), R, Cc, H * W, float(guidance), float(sqrt_beta_t),
                                         float(sqrt_alpha_t), L.cur_stream()), 'ddim_x0_score')
    _count()
    return scores, sums, x0


def sd_candidates(pivot: torch.Tensor, dirs: torch.Tensor, scale: torch.Tensor, lam: float, sqrt_e: float,
                  fresh: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SD eps_greedy / zero_order candidate noises: pivot fp32 [1,C,H,W]; dirs fp32 [N,C,H,W]; scale fp32 [N] (the rand(1)
    draws); fresh u8 [N].  cand = fresh ? dirs : pivot + ((dirs/||dirs|| * scale) * lam) * sqrt_e."""
    _chk_cuda(pivot, dirs, scale, fresh)
    _c(pivot, torch.float32), _c(dirs, torch.float32), _c(scale, torch.float32)
    if fresh is not None:
        _c(fresh, torch.uint8)
    N, E = dirs.shape[0], dirs[0].numel()
    if pivot.numel() != E or scale.numel() != N:
        raise RuntimeError('sd_candidates: shape mismatch')
    cand = torch.empty_like(dirs)
    L.check(L.lib().b200ns_sd_candidates(L.ptr(pivot), L.ptr(dirs), L.ptr(scale), L.ptr(fresh), L.ptr(cand), N, E,
                                         float(lam), float(sqrt_e), L.cur_stream()), 'sd_candidates')
    _count()
    return cand


def post_quant(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """1x1 conv C -> C on fp32 NCHW latents (AutoencoderKL.post_quant_conv)."""
    _chk_cuda(x, w, bias, out)
    _c(x, torch.float32), _c(w, torch.float32), _c(bias, torch.float32)
    B, Cc = x.shape[0], x.shape[1]
    out = torch.empty_like(x) if out is None else _c(out, torch.float32)
    L.check(L.lib().b200ns_post_quant(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(out), B, Cc, x[0, 0].numel(), L.cur_stream()),
            'post_quant')
    _count()
    return out


def image_sums(img: torch.Tensor, want_u8: bool = False):
    """img fp32 NHWC [B,H,W,C<=4] -> (integer channel sums int32 [B,4] of the uint8 quantisation, uint8 NCHW image or None)."""
    _chk_cuda(img)
    _c(img, torch.float32)
    B, H, W_, Cc = img.shape
    sums = torch.empty((B, 4), dtype=torch.int32, device=img.device)
    u8 = torch.empty((B, Cc, H, W_), dtype=torch.uint8, device=img.device) if want_u8 else None
    L.check(L.lib().b200ns_image_sums(L.ptr(img), L.ptr(sums), L.ptr(u8), B, Cc, H * W_, L.cur_stream()), 'image_sums')
    _count()
    return sums, u8


def pack_conv_up2(w: torch.Tensor, splits: Optional[Sequence[int]] = None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] fp32 -> bf16 [4, Cout, 4*Cin]: the four 2x2 phase kernels of `conv3x3(nearest_upsample_2x(x))`.
    Output pixel (2y+py, 2x+px) reads the low-res pixels (y+py-1+a, x+px-1+c), a, c in {0, 1}; the 3x3 taps that fall on the
    same source pixel are summed in fp32.  K order per phase = per input-channel split: (tap a*2+c, channel), like _pack_conv."""
    G = [[[0], [1, 2]], [[0, 1], [2]]]                        # [phase][tap] -> 3x3 kernel indices along one axis
    w = w.detach().float().cpu()
    out = []
    for py in range(2):
        for px in range(2):
            parts, c0 = [], 0
            for cs in (splits or [w.shape[1]]):
                ws = w[:, c0:c0 + cs]
                taps = [sum(ws[:, :, ky, kx] for ky in G[py][a] for kx in G[px][c]) for a in range(2) for c in range(2)]
                parts.append(torch.stack(taps, dim=1).reshape(w.shape[0], -1))          # [Cout, 4, cs] -> [Cout, 4*cs]
                c0 += cs
            out.append(torch.cat(parts, dim=1))
    return torch.stack(out).contiguous().to(ACT_DTYPE)


def interleave_geglu(t: torch.Tensor) -> torch.Tensor:
    """Rows [hidden (F) ; gate (F)] of a GEGLU projection (weight [2F, K] or bias [2F]) -> groups of
    [64 hidden | 64 gate] rows, the order the fused GEGLU epilogue of the GEMM expects.  F % 64 == 0."""
    F = t.shape[0] // 2
    h, g = t[:F].reshape(F // 64, 64, *t.shape[1:]), t[F:].reshape(F // 64, 64, *t.shape[1:])
    return torch.stack([h, g], dim=1).reshape(t.shape).contiguous()


# ------------------------------------------------------------------ plans
class Plan:
    """Ordered list of kernel launches (b200ns_plan).  Keeps every tensor it references alive."""

    def __init__(self):
        self._h = L.lib().b200ns_plan_create()
        self._keep: List[torch.Tensor] = []
        self.labels: List[str] = []
        self.flops: List[float] = []          # algorithmic FLOPs per op (0 for non-GEMM ops)
        self.kinds: List[str] = []

    def __del__(self):
        try:
            if self._h:
                L.lib().b200ns_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _k(self, *ts):
        for t in ts:
            if t is not None:
                _chk_cuda(t)
                self._keep.append(t)

    def __len__(self):
        return L.lib().b200ns_plan_size(self._h)

    def run(self, first: Optional[int] = None, last: Optional[int] = None):
        if first is None:
            if USE_TORCH_OPS:
                T.OPS.plan_run(int(self._h))
            else:
                L.check(L.lib().b200ns_plan_run(self._h, L.cur_stream()), 'plan_run')
            _count(len(self.labels))
        else:
            L.check(L.lib().b200ns_plan_run_range(self._h, first, last, L.cur_stream()), 'plan_run_range')
            _count(last - first)

    def set_lane(self, lane: int):
        """Ops added from now on go to graph branch `lane` (0 = main stream; 1..3 run in parallel when captured)."""
        L.check(L.lib().b200ns_plan_set_lane(self._h, lane), 'plan_set_lane')

    def set_pdl(self, mode: int):
        """Programmatic dependent launch for this plan (-1 process default, 0 off, 1 all kernels, 2 GroupNorm only)."""
        L.check(L.lib().b200ns_plan_set_pdl(self._h, int(mode)), 'plan_set_pdl')

    def instantiate_graph(self):
        """Capture the whole plan as one CUDA graph (static buffers): run() then costs one launch."""
        L.check(L.lib().b200ns_plan_instantiate_graph(self._h), 'plan_instantiate_graph')

    def run_timed(self):
        """Run op by op with CUDA events around each launch; returns per-op milliseconds."""
        n = len(self.labels)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            L.check(L.lib().b200ns_plan_run_range(self._h, i, i + 1, L.cur_stream()), 'plan_run_range')
            evs[i + 1].record()
        _count(n)
        torch.cuda.synchronize()
        return [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]

    def add_gemm(self, a: Sequence[torch.Tensor], segs: Sequence[Tuple[int, int, int, int]], w: torch.Tensor, N: int,
                 out: torch.Tensor, *, bias=None, residual=None, out_scale=1.0, gn_stats=None, reverse=False, a_stride=None,
                 label='gemm', alg_k=None, geglu=False, upsample2x=False, act=0, a_norm=None):
        """a: 1-3 NHWC bf16 tensors [B,H,W,C]; segs: (src, taps, cstart, cblocks); w: bf16 [Npad,Ktot].
        gn_stats: optional fp32 [M/64, N, 2] receiving per-channel (sum, sumsq) of the stored output.
        geglu: w / bias rows in groups of [64 hidden | 64 gate] (`interleave_geglu`); out is [.., N/2] = hidden * gelu(gate).
        upsample2x: out [B,2H,2W,N] = conv3x3(nearest_up2(a)) from the LOW-res a; w = pack_conv_up2(...) [4, Npad, Ktot];
        segs say taps = 9; gn_stats (optional) is fp32 [4*M/64, N, 2] over the high-res output."""
        d = L.GemmDesc()
        a_stride = list(a_stride) if a_stride is not None else [1] * len(a)
        B, H, W_, _ = a[0].shape
        H, W_ = H // a_stride[0], W_ // a_stride[0]           # output spatial dims (stride-2 sources are 2H x 2W)
        for i, t in enumerate(a):
            _c(t, ACT_DTYPE)
            if t.shape[1] != H * a_stride[i] or t.shape[2] != W_ * a_stride[i]:
                raise RuntimeError('gemm: source spatial dims do not match the output dims times the source stride')
            d.a_ptr[i] = L.ptr(t)
            d.a_channels[i] = t.shape[3]
            d.a_stride[i] = a_stride[i]
        d.n_seg = len(segs)
        for i, (src, taps, cstart, cblocks) in enumerate(segs):
            d.seg[i] = L.KSeg(src, taps, cstart, cblocks)
        d.batch, d.H, d.W = B, H, W_
        _c(w, ACT_DTYPE)
        up = 4 if upsample2x else 1
        if upsample2x:
            if w.dim() != 3 or w.shape[0] != 4 or tuple(out.shape[:3]) != (B, 2 * H, 2 * W_):
                raise RuntimeError('gemm(upsample2x): w must be [4, Npad, Ktot] (pack_conv_up2) and out [B, 2H, 2W, N]')
            d.w_ptr, d.N, d.Npad, d.Ktot = L.ptr(w), N, w.shape[1], w.shape[2]
        else:
            d.w_ptr, d.N, d.Npad, d.Ktot = L.ptr(w), N, w.shape[0], w.shape[1]
        d.upsample2x = int(bool(upsample2x))
        d.bias = L.ptr(bias)
        d.residual = L.ptr(residual)
        d.ld_res = residual.shape[-1] if residual is not None else 0
        d.out_scale = float(out_scale)
        d.out, d.ld_out = L.ptr(out), out.shape[-1]
        d.out_fp32 = 1 if out.dtype == torch.float32 else 0
        if gn_stats is not None:
            _c(gn_stats, torch.float32)
            if gn_stats.numel() != up * (B * H * W_ // 64) * N * 2:
                raise RuntimeError('gn_stats must be fp32 [M/64, N, 2]')
        d.gn_stats = L.ptr(gn_stats)
        d.reverse = int(reverse)
        d.geglu = int(bool(geglu))
        d.act = int(act)                                       # 1: quick_gelu(acc + bias) in the epilogue (CLIP MLP)
        if a_norm is not None:     # (mean_rstd fp32 [B, groups, 2], gamma, beta, groups): GroupNorm of A in the operand path
            mr, gam, bet, groups = a_norm
            d.xf_mean_rstd, d.xf_gamma, d.xf_beta = L.ptr(_c(mr, torch.float32)), L.ptr(_c(gam, torch.float32)), L.ptr(_c(bet, torch.float32))
            d.xf_groups = int(groups)
            self._k(mr, gam, bet)
        if geglu and out.shape[-1] * 2 != N:
            raise RuntimeError('gemm(geglu): out must have N/2 columns')
        self._k(*a, w, bias, residual, out, gn_stats)
        n_before = L.lib().b200ns_plan_size(self._h)
        L.check(L.lib().b200ns_plan_add_gemm(self._h, C.byref(d)), 'plan_add_gemm')
        n_after = L.lib().b200ns_plan_size(self._h)
        for j in range(n_before, n_after):             # one launch, or two over column slices (e.g. 320 = 192 + 128)
            cols = L.lib().b200ns_plan_gemm_cols(self._h, j)
            self.labels.append(label if n_after - n_before == 1 else f'{label}[{cols}]')
            self.kinds.append('gemm')
            self.flops.append(2.0 * B * H * W_ * cols * (alg_k if alg_k else w.shape[-1]))

    def add_gn_stats(self, x: Sequence[torch.Tensor], groups: int, partial: torch.Tensor, splits: int, *, pre_add=None,
                     b_emb=1, label='gn_stats'):
        d = L.GnStatsDesc()
        B, H, W_, _ = x[0].shape
        for i, t in enumerate(x):
            _c(t, ACT_DTYPE)
            d.x_ptr[i] = L.ptr(t)
            d.x_channels[i] = t.shape[3]
        d.batch, d.HW, d.groups = B, H * W_, groups
        d.pre_add = L.ptr(pre_add)
        d.ld_pre_add = pre_add.stride(0) if pre_add is not None else 0
        d.b_emb = b_emb
        _c(partial, torch.float64)
        d.partial, d.splits = L.ptr(partial), splits
        self._k(*x, pre_add, partial)
        L.check(L.lib().b200ns_plan_add_gn_stats(self._h, C.byref(d)), 'plan_add_gn_stats')
        self.labels.append(label)
        self.kinds.append('gn_stats')
        self.flops.append(0.0)

    def add_gn_finalize(self, stats: Sequence[torch.Tensor], channels: Sequence[int], batch: int, HW: int, groups: int,
                        eps: float, mean_rstd: torch.Tensor, *, pre_add=None, b_emb=1, label='gn_finalize', _desc_only=False):
        """stats: per-source fp32 [batch*HW/64, C_i, 2] written by the producing GEMMs -> mean_rstd fp32 [batch, groups, 2]."""
        d = L.GnFinalizeDesc()
        for i, (t, c) in enumerate(zip(stats, channels)):
            _c(t, torch.float32)
            d.stats_ptr[i] = L.ptr(t)
            d.x_channels[i] = c
        d.batch, d.HW, d.groups = batch, HW, groups
        d.pre_add = L.ptr(pre_add)
        d.ld_pre_add = pre_add.stride(0) if pre_add is not None else 0
        d.b_emb, d.eps = b_emb, float(eps)
        d.mean_rstd = L.ptr(_c(mean_rstd, torch.float32))
        self._k(*stats, pre_add, mean_rstd)
        if _desc_only:
            return d
        L.check(L.lib().b200ns_plan_add_gn_finalize(self._h, C.byref(d)), 'plan_add_gn_finalize')
        self.labels.append(label)
        self.kinds.append('gn_finalize')
        self.flops.append(0.0)

    def add_gn_apply(self, x: Sequence[torch.Tensor], groups: int, partial: Optional[torch.Tensor], splits: int, eps: float,
                     gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, *, pre_add=None, film_scale=None,
                     film_shift=None, b_emb=1, silu=True, resample=0, raw_out=None, mean_rstd=None, reverse=False,
                     label='gn_apply', _desc_only=False):
        d = L.GnApplyDesc()
        B, H, W_, _ = x[0].shape
        for i, t in enumerate(x):
            _c(t, ACT_DTYPE)
            d.x_ptr[i] = L.ptr(t)
            d.x_channels[i] = t.shape[3]
        d.batch, d.H, d.W, d.groups = B, H, W_, groups
        d.partial, d.splits, d.eps = L.ptr(partial), splits, float(eps)
        d.gamma, d.beta = L.ptr(_c(gamma, torch.float32)), L.ptr(_c(beta, torch.float32))
        d.pre_add = L.ptr(pre_add)
        d.ld_pre_add = pre_add.stride(0) if pre_add is not None else 0
        d.film_scale, d.film_shift = L.ptr(film_scale), L.ptr(film_shift)
        d.ld_film = film_scale.stride(0) if film_scale is not None else 0
        d.b_emb = b_emb
        d.silu, d.resample = int(silu), resample
        d.out, d.raw_out = L.ptr(_c(out, ACT_DTYPE)), L.ptr(raw_out)
        d.mean_rstd = L.ptr(mean_rstd)
        d.reverse = int(reverse)
        self._k(*x, partial, gamma, beta, pre_add, film_scale, film_shift, out, raw_out, mean_rstd)
        if _desc_only:
            return d
        L.check(L.lib().b200ns_plan_add_gn_apply(self._h, C.byref(d)), 'plan_add_gn_apply')
        self.labels.append(label)
        self.kinds.append('gn_apply')
        self.flops.append(0.0)

    def add_gn_norm(self, stats: Sequence[torch.Tensor], x: Sequence[torch.Tensor], groups: int, eps: float,
                    mean_rstd: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, *, pre_add=None,
                    film_scale=None, film_shift=None, b_emb=1, silu=True, resample=0, raw_out=None, reverse=False,
                    label='gn'):
        """GroupNorm from the producers' epilogue statistics: gn_finalize + gn_apply.  At the low-resolution levels
        (H*W <= 256; both kernels are latency-bound there) the two run as ONE cluster launch (gn_norm_cluster_kernel,
        bit-identical) when B200NS_GN_CLUSTER=1; measured no faster, so two launches by default."""
        B, H, W_, _ = x[0].shape
        chans = [t.shape[3] for t in x]
        fin = dict(pre_add=pre_add, b_emb=b_emb)
        app = dict(pre_add=pre_add, film_scale=film_scale, film_shift=film_shift, b_emb=b_emb, silu=silu, resample=resample,
                   raw_out=raw_out, mean_rstd=mean_rstd, reverse=reverse)
        if GN_CLUSTER and H * W_ <= 256 and (H * W_) % 64 == 0 and sum(chans) <= 2048:
            f = self.add_gn_finalize(stats, chans, B, H * W_, groups, eps, mean_rstd, _desc_only=True, **fin)
            a = self.add_gn_apply(x, groups, None, 1, eps, gamma, beta, out, _desc_only=True, **app)
            L.check(L.lib().b200ns_plan_add_gn_norm(self._h, C.byref(f), C.byref(a)), 'plan_add_gn_norm')
            self.labels.append(f'{label}.norm')
            self.kinds.append('gn_norm')
            self.flops.append(0.0)
            return
        self.add_gn_finalize(stats, chans, B, H * W_, groups, eps, mean_rstd, label=f'{label}.finalize', **fin)
        self.add_gn_apply(x, groups, None, 1, eps, gamma, beta, out, label=f'{label}.apply', **app)

    def add_attention(self, qk: torch.Tensor, k_col0: int, vt: Optional[torch.Tensor], out: torch.Tensor, batch: int,
                      heads: int, Lseq: int, label='attention', v_col0: int = 0, head_dim: int = 64, reverse=False,
                      scale: float = 0.0, kv: Optional[torch.Tensor] = None, kv_rows: int = 0, kv_len: int = 0, kv_div: int = 1,
                      kv_ld: int = 0):
        """qk: [batch*L, ld] with Q at col head*64, K at k_col0 + head*64; V either transposed in `vt`
        ([batch*heads*64, L]) or (vt=None) row-major in `qk` at v_col0 + head*64."""
        d = L.AttnDesc()
        d.qk, d.ld_qk, d.k_col0 = L.ptr(_c(qk, ACT_DTYPE)), qk.shape[-1], k_col0
        d.vt = L.ptr(vt)
        d.v_col0 = v_col0
        d.head_dim = head_dim
        d.reverse = int(reverse)
        d.scale = float(scale)
        if kv is not None:         # cross-attention: K/V of the (few) contexts, [kv_batch * kv_rows, ld_kv]
            if kv_ld:              # a column window [K | V] of a wider row-major matrix (e.g. of a fused QKV projection)
                if kv.dtype != ACT_DTYPE or kv.stride(-1) != 1 or kv.stride(0) != kv_ld:
                    raise RuntimeError('attention: kv window must be a column slice of a contiguous activation matrix')
                d.kv, d.ld_kv = L.ptr(kv), kv_ld
            else:
                d.kv, d.ld_kv = L.ptr(_c(kv, ACT_DTYPE)), kv.shape[-1]
            d.kv_batch, d.kv_rows, d.kv_len, d.kv_div = kv.numel() // (kv.shape[-1] * kv_rows), kv_rows, kv_len, kv_div
            self._keep.append(kv)
        d.out, d.ld_out = L.ptr(_c(out, ACT_DTYPE)), out.shape[-1]
        d.batch, d.heads, d.L = batch, heads, Lseq
        self._k(qk, vt, out)
        L.check(L.lib().b200ns_plan_add_attention(self._h, C.byref(d)), 'plan_add_attention')
        self.labels.append(label)
        self.kinds.append('attention')
        self.flops.append(4.0 * batch * heads * Lseq * (kv_rows if kv is not None else Lseq) * head_dim)

    def add_linear(self, x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias=None, add=None, act=0,
                   label='linear'):
        d = L.LinearDesc()
        _c(w, torch.float32)
        d.x, d.rows, d.K, d.ld_x = L.ptr(x), x.shape[0], w.shape[1], x.stride(0)
        d.w, d.bias, d.add = L.ptr(w), L.ptr(bias), L.ptr(add)
        d.ld_add = add.stride(0) if add is not None else 0
        d.N, d.act = w.shape[0], act
        d.out, d.ld_out = L.ptr(out), out.stride(0)
        self._k(x, w, bias, add, out)
        L.check(L.lib().b200ns_plan_add_linear(self._h, C.byref(d)), 'plan_add_linear')
        self.labels.append(label)
        self.kinds.append('linear')
        self.flops.append(0.0)

    def add_layernorm(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, eps: float = 1e-5,
                      label='layernorm'):
        """nn.LayerNorm over the last dim of a bf16 tensor [..., C]."""
        Cc = x.shape[-1]
        self._k(x, gamma, beta, out)
        L.check(L.lib().b200ns_plan_add_layernorm(self._h, L.ptr(_c(x, ACT_DTYPE)), L.ptr(_c(gamma, torch.float32)),
                                                  L.ptr(_c(beta, torch.float32)), L.ptr(_c(out, ACT_DTYPE)),
                                                  x.numel() // Cc, Cc, float(eps)), 'plan_add_layernorm')
        self._misc('layernorm', label)

    def add_geglu(self, x: torch.Tensor, out: torch.Tensor, label='geglu'):
        """x bf16 [..., 2F] = [hidden | gate] -> out bf16 [..., F] = hidden * gelu(gate)."""
        F_ = out.shape[-1]
        if x.shape[-1] != 2 * F_:
            raise RuntimeError('geglu: input must be twice as wide as the output')
        self._k(x, out)
        L.check(L.lib().b200ns_plan_add_geglu(self._h, L.ptr(_c(x, ACT_DTYPE)), L.ptr(_c(out, ACT_DTYPE)),
                                              out.numel() // F_, F_), 'plan_add_geglu')
        self._misc('geglu', label)

    def add_softmax_rows(self, S: torch.Tensor, P: torch.Tensor, scale: float, label='softmax'):
        """S fp32 [rows, L] -> P bf16 [rows, L] = softmax(scale * S) over the last axis."""
        self._k(S, P)
        rows, L_ = S.shape[-2] * (S.numel() // (S.shape[-1] * S.shape[-2])), S.shape[-1]
        L.check(L.lib().b200ns_plan_add_softmax_rows(self._h, L.ptr(_c(S, torch.float32)), L.ptr(_c(P, ACT_DTYPE)), rows, L_,
                                                     float(scale)), 'plan_add_softmax_rows')
        self._misc('softmax_rows', label)

    def add_upsample2x(self, x: torch.Tensor, out: torch.Tensor, label='upsample2x'):
        B, H, W_, Cc = x.shape
        self._k(x, out)
        L.check(L.lib().b200ns_plan_add_upsample2x(self._h, L.ptr(_c(x, ACT_DTYPE)), L.ptr(_c(out, ACT_DTYPE)),
                                                   B, H, W_, Cc), 'plan_add_upsample2x')
        self._misc('upsample2x', label)

    def add_im2col(self, x: torch.Tensor, out: torch.Tensor, label='im2col'):
        d = L.Im2colDesc()
        B, Cc, H, W_ = x.shape
        d.x, d.out = L.ptr(_c(x, torch.float32)), L.ptr(_c(out, ACT_DTYPE))
        d.batch, d.C, d.H, d.W = B, Cc, H, W_
        self._k(x, out)
        L.check(L.lib().b200ns_plan_add_im2col(self._h, C.byref(d)), 'plan_add_im2col')
        self.labels.append(label)
        self.kinds.append('im2col')
        self.flops.append(0.0)

    # ---- split-fp16 "precise" ops (csrc/precise.cuh): near-tie re-scoring of a search round's contenders
    def add_gemm_prec(self, a: Sequence[torch.Tensor], segs: Sequence[Tuple[int, int, int, int]], w: torch.Tensor, N: int,
                      out: torch.Tensor, *, acc_scale: float, bias=None, residual=None, out_scale=1.0, label='gemm_prec',
                      flops: float = 0.0, splits: int = 1, bn: int = 0, partial: Optional[torch.Tensor] = None,
                      ticket: Optional[torch.Tensor] = None):
        """a: 1-3 split-half NHWC tensors [B,H,W,2C] (hi plane | lo plane); segs: (src, taps, cstart, cblocks) over the C
        LOGICAL channels (= the hi plane); w: half, K-block-major [2*K/64, Npad, 64] = per K block the Whi tile, then the Wlo
        tile (precise.pack_split; pre-scaled by 1/acc_scale); out: split half [B,H,W,2N] or fp32 [..., N]; residual: split
        half [B,H,W,2N].  Per K block the kernel issues hi x Whi + lo x Whi + hi x Wlo."""
        d = L.GemmDesc()
        B, H, W_, _ = a[0].shape
        for i, t in enumerate(a):
            _c(t, torch.float16)
            if t.shape[1] != H or t.shape[2] != W_:
                raise RuntimeError('gemm_prec: sources must share the spatial dims')
            d.a_ptr[i] = L.ptr(t)
            d.a_channels[i] = t.shape[3]
            d.a_stride[i] = 1
        d.n_seg = len(segs)
        for i, (src, taps, cstart, cblocks) in enumerate(segs):
            d.seg[i] = L.KSeg(src, taps, cstart, cblocks)
        d.batch, d.H, d.W = B, H, W_
        _c(w, torch.float16)
        if w.dim() != 3 or w.shape[2] != 64 or w.shape[0] % 2:
            raise RuntimeError('gemm_prec: w must be K-block-major [2*K/64, Npad, 64] (precise.pack_split)')
        d.w_ptr, d.N, d.Npad, d.Ktot = L.ptr(w), N, w.shape[1], (w.shape[0] // 2) * 64
        d.bias = L.ptr(bias)
        d.out_scale = float(out_scale)
        d.out = L.ptr(out)
        d.ld_out = out.shape[-1]
        d.out_fp32 = 1 if out.dtype == torch.float32 else 0
        if not d.out_fp32:
            _c(out, torch.float16)
            if out.shape[-1] != 2 * N:
                raise RuntimeError('gemm_prec: split output must have 2N columns')
            d.out_lo_off = N
        if residual is not None:
            _c(residual, torch.float16)
            d.residual, d.ld_res, d.res_lo_off = L.ptr(residual), residual.shape[-1], residual.shape[-1] // 2
        d.prec, d.acc_scale = 1, float(acc_scale)
        d.prec_splits, d.prec_bn = int(splits), int(bn)
        if splits > 1:
            # fp32 partial tiles of the K slices: [splits, ceil(M/128)*128, Npad]
            need = splits * ((B * H * W_ + 127) // 128) * 128 * w.shape[1]
            if partial is None or partial.dtype != torch.float32 or partial.numel() < need:
                raise RuntimeError('gemm_prec: split-K needs an fp32 `partial` workspace of splits*Mpad*Npad elements')
            d.prec_partial = L.ptr(partial)
            if ticket is not None:               # in-kernel finish (the last K slice of a tile adds all of them)
                d.prec_ticket, d.prec_ticket_len = L.ptr(_c(ticket, torch.int32)), ticket.numel()
        self._k(*a, w, bias, residual, out, partial, ticket)
        L.check(L.lib().b200ns_plan_add_gemm(self._h, C.byref(d)), 'plan_add_gemm(prec)')
        self.labels.append(label)
        self.kinds.append('gemm_prec')
        self.flops.append(flops)

    def _gn_prec_desc(self, x, groups, eps, mean_rstd, gamma=None, beta=None, out=None, *, pre_add=None, film_scale=None,
                      film_shift=None, b_emb=1, silu=False, resample=0, raw_out=None):
        d = L.GnPrecDesc()
        B, H, W_, _ = x[0].shape
        # statistics scratch, shared by every GroupNorm of the plan (ops run in stream order): fp64 partial sums per
        # (sample, pixel split, group) and the per-sample arrival counters (zeroed once; the kernel resets them)
        scr = getattr(self, '_gn_prec_scratch', None)
        if scr is None or scr[0].shape[0] < B:
            dev = x[0].device
            scr = (torch.empty(B, 64, 64, 2, dtype=torch.float64, device=dev), torch.zeros(B, dtype=torch.int32, device=dev))
            self._gn_prec_scratch = scr
        if groups > 64:
            raise RuntimeError('gn_prec: at most 64 groups')
        d.partial, d.ticket = L.ptr(scr[0]), L.ptr(scr[1])
        self._k(*scr)
        for i, t in enumerate(x):
            _c(t, torch.float16)
            d.x_ptr[i] = L.ptr(t)
            d.x_channels[i] = t.shape[3] // 2
        d.batch, d.H, d.W, d.groups, d.eps = B, H, W_, groups, float(eps)
        d.gamma = L.ptr(_c(gamma, torch.float32)) if gamma is not None else None
        d.beta = L.ptr(_c(beta, torch.float32)) if beta is not None else None
        d.pre_add = L.ptr(pre_add)
        d.ld_pre_add = pre_add.stride(0) if pre_add is not None else 0
        d.film_scale, d.film_shift = L.ptr(film_scale), L.ptr(film_shift)
        d.ld_film = film_scale.stride(0) if film_scale is not None else 0
        d.b_emb, d.silu, d.resample = b_emb, int(silu), resample
        d.out = L.ptr(_c(out, torch.float16)) if out is not None else None
        d.raw_out = L.ptr(raw_out)
        d.mean_rstd = L.ptr(_c(mean_rstd, torch.float32))
        self._k(*x, gamma, beta, pre_add, film_scale, film_shift, out, raw_out, mean_rstd)
        return d

    def add_gn_prec(self, x: Sequence[torch.Tensor], groups: int, eps: float, mean_rstd: torch.Tensor, gamma, beta, out, *,
                    pre_add=None, film_scale=None, film_shift=None, b_emb=1, silu=True, resample=0, raw_out=None, label='gn_prec'):
        """GroupNorm (+FiLM, +SiLU, +2x resample) over split-half NHWC tensors [B,H,W,2C_i]: one statistics launch (fp64)
        and one apply launch (fp32, exact expf / division)."""
        d = self._gn_prec_desc(x, groups, eps, mean_rstd, gamma, beta, out, pre_add=pre_add, film_scale=film_scale,
                               film_shift=film_shift, b_emb=b_emb, silu=silu, resample=resample, raw_out=raw_out)
        L.check(L.lib().b200ns_plan_add_gn_stats_prec(self._h, C.byref(d)), 'plan_add_gn_stats_prec')
        self._misc('gn_stats_prec', f'{label}.stats')
        L.check(L.lib().b200ns_plan_add_gn_apply_prec(self._h, C.byref(d)), 'plan_add_gn_apply_prec')
        self._misc('gn_apply_prec', f'{label}.apply')

    def add_attention_prec(self, qkv: torch.Tensor, out: torch.Tensor, batch: int, heads: int, Lseq: int, Cc: int,
                           label='attention_prec'):
        """qkv split half [batch*L, 2*3C] ([Q|K|V] head-major inside each plane); out split half [batch*L, 2C]."""
        d = L.AttnPrecDesc()
        d.qkv, d.ld, d.lo_off, d.k_col0, d.v_col0 = L.ptr(_c(qkv, torch.float16)), qkv.shape[-1], 3 * Cc, Cc, 2 * Cc
        d.out, d.ld_out, d.out_lo_off = L.ptr(_c(out, torch.float16)), out.shape[-1], Cc
        d.batch, d.heads, d.L, d.scale = batch, heads, Lseq, 0.0
        self._k(qkv, out)
        L.check(L.lib().b200ns_plan_add_attention_prec(self._h, C.byref(d)), 'plan_add_attention_prec')
        self._misc('attention_prec', label)

    def add_im2col_prec(self, x: torch.Tensor, out: torch.Tensor, label='im2col_prec'):
        d = L.Im2colDesc()
        B, Cc, H, W_ = x.shape
        d.x, d.out = L.ptr(_c(x, torch.float32)), L.ptr(_c(out, torch.float16))
        d.batch, d.C, d.H, d.W = B, Cc, H, W_
        self._k(x, out)
        L.check(L.lib().b200ns_plan_add_im2col_prec(self._h, C.byref(d)), 'plan_add_im2col_prec')
        self._misc('im2col_prec', label)

    def _misc(self, kind, label):
        self.labels.append(label)
        self.kinds.append(kind)
        self.flops.append(0.0)

    def add_clip_preprocess(self, img: torch.Tensor, tmp: torch.Tensor, patches: torch.Tensor, hb, hk, vb, vk, lut, S: int, P: int,
                            Lp: int, label='clip_preprocess'):
        """uint8 [B,3,H,W] -> the patch-embedding GEMM's A matrix (activation type) [B*Lp, Kp]: Pillow's bicubic resize + centre
        crop + rescale/normalise of CLIPImageProcessor, bit-exact (csrc/clip.cuh).  hb/vb int32 [S,2], hk/vk int32 [S,ks]."""
        d = L.ClipPreprocessDesc()
        B, _, H, W_ = img.shape
        d.img, d.tmp, d.patches = L.ptr(_c(img, torch.uint8)), L.ptr(_c(tmp, torch.uint8)), L.ptr(_c(patches, ACT_DTYPE))
        if tmp.numel() < B * 3 * H * S or patches.shape[0] != B * Lp or img.shape[1] != 3:
            raise RuntimeError('clip_preprocess: tmp must hold [B,3,H,S] bytes and patches must be [B*Lp, Kp]')
        for t in (hb, hk, vb, vk):
            _c(t, torch.int32)
        if tuple(hb.shape) != (S, 2) or tuple(vb.shape) != (S, 2) or hk.shape[0] != S or vk.shape[0] != S:
            raise RuntimeError('clip_preprocess: bounds must be [S,2] and coefficients [S,ks]')
        d.h_bounds, d.h_coeffs, d.v_bounds, d.v_coeffs = L.ptr(hb), L.ptr(hk), L.ptr(vb), L.ptr(vk)
        d.lut = L.ptr(_c(lut, torch.float32))
        d.batch, d.H, d.W, d.S, d.P, d.Lp, d.Kp, d.hks, d.vks = B, H, W_, S, P, Lp, patches.shape[1], hk.shape[1], vk.shape[1]
        self._k(img, tmp, patches, hb, hk, vb, vk, lut)
        L.check(L.lib().b200ns_plan_add_clip_preprocess(self._h, C.byref(d)), 'plan_add_clip_preprocess')
        self._misc('clip_preprocess', label)

    def add_clip_pool_ln(self, x: torch.Tensor, row_stride: int, gamma, beta, out: torch.Tensor, batch: int, eps: float = 1e-5,
                         label='clip_pool_ln'):
        self._k(x, gamma, beta, out)
        L.check(L.lib().b200ns_plan_add_clip_pool_ln(self._h, L.ptr(_c(x, ACT_DTYPE)), row_stride, L.ptr(_c(gamma, torch.float32)),
                                                     L.ptr(_c(beta, torch.float32)), L.ptr(_c(out, torch.float32)), batch,
                                                     out.shape[-1], float(eps)), 'plan_add_clip_pool_ln')
        self._misc('clip_pool_ln', label)

    def add_clip_cosine(self, image_embeds: torch.Tensor, text_embeds: torch.Tensor, score: torch.Tensor, label='clip_cosine'):
        self._k(image_embeds, text_embeds, score)
        B, D = image_embeds.shape
        L.check(L.lib().b200ns_plan_add_clip_cosine(self._h, L.ptr(_c(image_embeds, torch.float32)),
                                                    L.ptr(_c(text_embeds, torch.float32)), text_embeds.shape[0],
                                                    L.ptr(_c(score, torch.float32)), B, D), 'plan_add_clip_cosine')
        self._misc('clip_cosine', label)

    def add_u8_to_f32(self, src: torch.Tensor, dst: torch.Tensor, label='u8_to_f32'):
        self._k(src, dst)
        L.check(L.lib().b200ns_plan_add_u8_to_f32(self._h, L.ptr(_c(src, torch.uint8)), L.ptr(_c(dst, torch.float32)),
                                                  src.numel()), 'plan_add_u8_to_f32')
        self._misc('misc', label)

    def add_pool_tokens(self, act, pos, tok, tok0, batch, T, Cc, label='pool_tokens'):
        self._k(act, pos, tok, tok0)
        L.check(L.lib().b200ns_plan_add_pool_tokens(self._h, L.ptr(_c(act, ACT_DTYPE)), L.ptr(_c(pos, torch.float32)),
                                                    L.ptr(_c(tok, ACT_DTYPE)), L.ptr(_c(tok0, torch.float32)), batch, T,
                                                    Cc), 'plan_add_pool_tokens')
        self._misc('misc', label)

    def add_pool_attention(self, qkv0, kv, out, batch, T, Cc, label='pool_attention'):
        self._k(qkv0, kv, out)
        L.check(L.lib().b200ns_plan_add_pool_attention(self._h, L.ptr(_c(qkv0, torch.float32)), L.ptr(_c(kv, ACT_DTYPE)),
                                                       L.ptr(_c(out, torch.float32)), batch, T, Cc), 'plan_add_pool_attention')
        self._misc('misc', label)

    def add_softmax_gather(self, logits, target, scores, label='softmax_gather'):
        self._k(logits, target, scores)
        L.check(L.lib().b200ns_plan_add_softmax_gather(self._h, L.ptr(_c(logits, torch.float32)),
                                                       L.ptr(_c(target, torch.int64)), L.ptr(_c(scores, torch.float32)),
                                                       logits.shape[0], logits.shape[1]), 'plan_add_softmax_gather')
        self._misc('misc', label)

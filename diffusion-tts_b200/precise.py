"""fp32-faithful re-evaluation engine for near-tie contenders (SURVEY.md 7 hard part 1, option b).

The reference evaluates the denoiser in fp32 (edm/training/networks.py:655-667) and picks `scores.argmax(dim=0)`
(edm/main.py:842); the bf16 tensor-core engine (`unet.py`) carries ~1e-4 of score noise, enough to flip the argmax when
the two best candidates are closer than that.  `PreciseUNetEngine` evaluates the SAME network (same state dict, same
plan structure as `unet.ForwardPlan`) for the handful of contenders in "split fp16": every activation and every conv
weight is a pair of IEEE halves (hi, lo = half(v - hi)), the GEMMs run on the same tcgen05 implicit-GEMM main loop over
three K segments ([hi|lo] x [Whi|Whi] + [hi] x [Wlo], fp32 accumulation in TMEM), GroupNorm statistics in fp64,
SiLU / softmax with IEEE expf and division, attention in fp32 on the FMA pipe (csrc/precise.cuh).  Measured against the
reference's fp32 forward: see tests/test_precise_gpu.py (relative L2 of F_x ~1e-6, the level of fp32 summation-order
noise between two fp32 implementations).

Only what the path needs: DhariwalUNet / SongUNet blocks with head_dim-64 attention (ADM).  Everything is batch-position
and batch-size invariant (order-fixed reductions), like the bf16 engine.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .ops import Plan
from .unet import Block, UNetEngine, _groups


def split_half(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    return hi, lo


def pack_split(blocks: Sequence[torch.Tensor], n_pad: Optional[int] = None):
    """blocks: fp32 [N, taps, C_i] weight blocks that share one accumulator (K-concatenated).
    Returns (half [2*K/64, Npad, 64], acc_scale, segs): K-block-major, per logical K block of 64 channels the Whi tile followed
    by the Wlo tile -- the B operands of `hi x Whi + lo x Whi + hi x Wlo` -- weights pre-multiplied by a power of two so that
    |W| <= 1024 (Wlo stays a normal half for all but the smallest weights); segs = per block (taps, C/64) over the LOGICAL
    channels (the hi plane of the activation; its lo plane lies C further right)."""
    N = blocks[0].shape[0]
    amax = max(float(b.abs().max()) for b in blocks)
    k = int(math.floor(10 - math.log2(amax))) if amax > 0 else 0
    k = max(-14, min(24, k))
    scale = 2.0 ** k
    parts, segs = [], []
    for b in blocks:
        taps, C = b.shape[1], b.shape[2]
        if C % 64:
            raise ValueError('split GEMM: channel blocks must be multiples of 64')
        hi, lo = split_half(b.to(torch.float32) * scale)
        if os.environ.get('B200NS_PREC_NOLO') == '1':          # experiment: plain fp16 weights (see b200ns_debug_prec_nolo)
            lo = torch.zeros_like(lo)
        parts.append(torch.stack([hi.reshape(N, taps * C), lo.reshape(N, taps * C)]))          # [2, N, K_i]
        segs.append((taps, C // 64))
    w = torch.cat(parts, dim=2)                                                                 # [2, N, K]
    if n_pad is not None and n_pad > N:
        w = torch.cat([w, torch.zeros(2, n_pad - N, w.shape[2], dtype=w.dtype)], dim=1)
    # K-block-major [K/64][hi, lo][Npad][64]: every operand tile of a K block is one contiguous run of HBM
    w = w.reshape(2, w.shape[1], w.shape[2] // 64, 64).permute(2, 0, 1, 3).contiguous()
    return w.reshape(-1, w.shape[2], 64), 1.0 / scale, segs


def flat_segs(segs) -> List[Tuple[int, int, int, int]]:
    """pack_split's per-block (taps, cblocks) -> the (src, taps, cstart, cblocks) list of `Plan.add_gemm_prec`, block i
    reading activation source i from its first channel."""
    return [(i, taps, 0, cblocks) for i, (taps, cblocks) in enumerate(segs)]


def split_k_policy(hw: int, n_pad: int, nkb: int, sms: int = 148) -> Tuple[int, int]:
    """(N tile width, K slices) of a precise GEMM.  Contender batches are tiny, so a layer is cut until ONE sample's worth
    of work items (M tiles x N tiles x K slices) fills the SMs; both numbers depend on the layer only -- never on the
    batch -- so a sample's bits do not depend on how many contenders share its launch."""
    if n_pad % 64:
        return 0, 1                               # the 3-channel output conv (Npad = 16): cost model, no split
    # (measured: forcing 192-wide tiles with up to 64 slices -- every SM busy at batch 1 -- was 10-20 % SLOWER than this
    # policy with the C side's cost model choosing the width: more partial-tile traffic and finishing work)
    bn = 128 if n_pad % 128 == 0 else 64
    tiles = max(1, hw // 128) * (n_pad // bn)
    splits = max(1, min(-(-sms // tiles), (3 * nkb) // 8, 32))     # nkb = logical K blocks (3 MMAs each)
    pref = int(os.environ.get('B200NS_PREC_BN_PREF', '0'))         # experiment: force a tile width where it divides
    if pref and n_pad % pref == 0:
        return pref, splits
    return 0, splits


def _conv_block(w: torch.Tensor, c0: int = 0, c1: Optional[int] = None) -> torch.Tensor:
    """[Cout, Cin, k, k] fp32 -> [Cout, k*k, Cin[c0:c1]] (tap-major, channel-minor: the A operand's K order)."""
    w = w.detach().float().cpu()
    c1 = w.shape[1] if c1 is None else c1
    return w[:, c0:c1].permute(0, 2, 3, 1).reshape(w.shape[0], w.shape[2] * w.shape[3], c1 - c0).contiguous()


class PreciseForwardPlan:
    def __init__(self, eng: 'PreciseUNetEngine', B: int, b_emb: int):
        cfg, dev = eng.cfg, eng.device
        self.B, self.b_emb = B, b_emb
        H = cfg.img_resolution
        f32 = dict(device=dev, dtype=torch.float32)
        self.x_in = torch.zeros(B, cfg.in_channels, H, H, **f32)
        self.emb_in = torch.zeros(b_emb, cfg.noise_channels, **f32)
        self.labels = torch.zeros(b_emb, max(cfg.label_dim, 1), **f32)
        self.out = torch.empty(B, H, H, cfg.out_channels, **f32)
        self.plan = Plan()
        self._scratch: Dict[str, torch.Tensor] = {}
        self.block_out: Dict[str, torch.Tensor] = {}
        self._eps = cfg.eps
        self._build(eng)
        if os.environ.get('B200NS_PDL') is None and os.environ.get('B200NS_PREC_PDL', '1') != '0':
            # contender batches are a handful of rows: ~470 short, latency-bound launches per forward -> overlap every kernel's
            # prologue with its predecessor's tail (programmatic dependent launch; all precise kernels call pdl_wait())
            self.plan.set_pdl(1)
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            self.plan.instantiate_graph()
            self.plan.run()               # the first launch uploads the ~470-node graph (milliseconds): pay it here, at build
            torch.cuda.synchronize(dev)   # time, not inside the first search round that happens to have this many contenders

    def _act(self, key: str, B, H, W, C) -> torch.Tensor:
        """Split-half activation [B,H,W,2C] in a named scratch buffer."""
        n = B * H * W * 2 * C
        t = self._scratch.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(n, device=self.x_in.device, dtype=torch.float16)
            self._scratch[key] = t
        return t[:n].view(B, H, W, 2 * C)

    def _new(self, B, H, W, C) -> torch.Tensor:
        return torch.empty(B, H, W, 2 * C, device=self.x_in.device, dtype=torch.float16)

    def _gn(self, xs, C, gamma, beta, out, *, silu=True, resample=0, raw_out=None, film=None, pre_add=None, label=''):
        g = _groups(C)
        mr = torch.empty(self.B, g, 2, device=self.x_in.device, dtype=torch.float32)
        self.plan.add_gn_prec(xs, g, self._eps, mr, gamma, beta, out, pre_add=pre_add,
                              film_scale=film[0] if film else None, film_shift=film[1] if film else None,
                              b_emb=self.b_emb, silu=silu, resample=resample, raw_out=raw_out, label=label)

    def _gemm(self, srcs, key, N, out, *, bias, residual=None, out_scale=1.0, label=''):
        w, acc_scale, segs = self._eng.w[key]
        flat = flat_segs(segs)
        B, H, W_, _ = srcs[0].shape
        nkb = w.shape[0] // 2                                       # logical K blocks: [nkb][Whi, Wlo][Npad][64]
        bn, splits = split_k_policy(H * W_, w.shape[1], nkb)
        partial = None
        if splits > 1:
            need = splits * ((B * H * W_ + 127) // 128) * 128 * w.shape[1]
            partial = self._scratch.get('splitk')
            if partial is None or partial.numel() < need:
                partial = torch.empty(need, device=self.x_in.device, dtype=torch.float32)
                self._scratch['splitk'] = partial
        self.plan.add_gemm_prec(srcs, flat, w, N, out, acc_scale=acc_scale, bias=bias, residual=residual,
                                out_scale=out_scale, label=label, flops=2.0 * B * H * W_ * N * 3 * nkb * 64,
                                splits=splits, bn=bn, partial=partial)

    def _build(self, eng: 'PreciseUNetEngine'):
        self._eng = eng
        cfg, P, W_ = eng.cfg, self.plan, eng.base.w
        b_emb = self.b_emb
        adm = cfg.model_type == 'DhariwalUNet'
        dev = self.x_in.device
        f32 = dict(device=dev, dtype=torch.float32)
        E = cfg.emb_channels
        # ---- embedding network: fp32 linear kernels, identical to the bf16 engine's (unet.py:_build)
        t0 = torch.empty(b_emb, E, **f32)
        self.emb = torch.empty(b_emb, E, **f32)
        if adm:
            P.add_linear(self.emb_in, W_['map_layer0.weight'], t0, bias=W_['map_layer0.bias'], act=1, label='map_layer0')
            lab = None
            if cfg.label_dim:
                lab = torch.empty(b_emb, E, **f32)
                P.add_linear(self.labels, W_['map_label.weight'], lab, label='map_label')
            P.add_linear(t0, W_['map_layer1.weight'], self.emb, bias=W_['map_layer1.bias'], add=lab, act=1, label='map_layer1')
        else:
            src = self.emb_in
            if cfg.label_dim:
                src = torch.empty(b_emb, cfg.noise_channels, **f32)
                P.add_linear(self.labels, W_['map_label.weight_scaled'], src, bias=W_['map_label.bias'], add=self.emb_in,
                             label='map_label')
            P.add_linear(src, W_['map_layer0.weight'], t0, bias=W_['map_layer0.bias'], act=1, label='map_layer0')
            P.add_linear(t0, W_['map_layer1.weight'], self.emb, bias=W_['map_layer1.bias'], act=1, label='map_layer1')
        self.film = torch.empty(b_emb, eng.base.affine_total, **f32)
        P.add_linear(self.emb, W_['affine_all.weight'], self.film, bias=W_['affine_all.bias'], label='affine_all')

        st = dict(x=None, skips=[], aux_in=None)
        for blk in cfg.enc:
            self._item('enc', blk, st)
        for blk in cfg.dec:
            self._item('dec', blk, st)
        if adm:
            self._item('out', None, st)

    def _item(self, kind: str, blk: Optional[Block], st: dict):
        eng = self._eng
        cfg, P, W_ = eng.cfg, self.plan, eng.base.w
        B = self.B
        if kind == 'out':                                    # out_norm + SiLU + out_conv (networks.py:460)
            H = cfg.img_resolution
            x = st['x']
            C = x.shape[3] // 2
            a = self._act('a0', B, H, H, C)
            self._gn([x], C, W_['out_norm.weight'], W_['out_norm.bias'], a, silu=True, label='out_norm')
            self._gemm([a], 'out_conv', cfg.out_channels, self.out, bias=W_['out_conv.b'], label='out_conv')
        elif kind == 'enc':
            if blk.kind == 'conv':
                H = blk.res
                col = self._act('col', B, H, H, 64)
                P.add_im2col_prec(self.x_in, col, label=f'{blk.name}.im2col')
                x = self._new(B, H, H, blk.cout)
                self._gemm([col], blk.name, blk.cout, x, bias=W_[f'{blk.name}.b'], label=blk.name)
                st['x'] = x
            else:
                st['x'] = self._block(blk, [st['x']])
            self.block_out[blk.name] = st['x']
            st['skips'].append(st['x'])
        elif blk.kind == 'aux_norm':
            H = blk.res
            st['aux_in'] = self._act('a0', B, H, H, blk.cin)
            self._gn([st['x']], blk.cin, W_[f'{blk.name}.weight'], W_[f'{blk.name}.bias'], st['aux_in'], silu=True, label=blk.name)
        elif blk.kind == 'aux_conv':
            self._gemm([st['aux_in']], blk.name, blk.cout, self.out, bias=W_[f'{blk.name}.b'], label=blk.name)
        else:
            xs = [st['x']]
            if st['x'].shape[3] // 2 != blk.cin:
                xs.append(st['skips'].pop())
                assert (xs[0].shape[3] + xs[1].shape[3]) // 2 == blk.cin
            st['x'] = self._block(blk, xs)
            self.block_out[blk.name] = st['x']

    def _block(self, blk: Block, xs: List[torch.Tensor]) -> torch.Tensor:
        """UNetBlock.forward (networks.py:166-187), same op sequence as unet.ForwardPlan._block without the fusions."""
        eng = self._eng
        cfg, P, W_ = eng.cfg, self.plan, eng.base.w
        B, n = self.B, blk.name
        Hin, Ho = xs[0].shape[1], blk.res
        cin, cout = blk.cin, blk.cout
        resample = 1 if blk.up else (2 if blk.down else 0)
        xr = self._act('xr', B, Ho, Ho, cin) if resample else None
        a0 = self._act('a0', B, Ho, Ho, cin)
        self._gn(xs, cin, W_[f'{n}.norm0.weight'], W_[f'{n}.norm0.bias'], a0, silu=True, resample=resample, raw_out=xr,
                 label=f'{n}.norm0')
        h = self._act('h', B, Ho, Ho, cout)
        self._gemm([a0], f'{n}.conv0', cout, h, bias=W_[f'{n}.conv0.b'], label=f'{n}.conv0')
        a1 = self._act('a1', B, Ho, Ho, cout)
        off = eng.base.affine_off[n]
        if cfg.adaptive_scale:
            film = (self.film[:, off:off + cout], self.film[:, off + cout:off + 2 * cout])
            self._gn([h], cout, W_[f'{n}.norm1.weight'], W_[f'{n}.norm1.bias'], a1, film=film, label=f'{n}.norm1')
        else:
            self._gn([h], cout, W_[f'{n}.norm1.weight'], W_[f'{n}.norm1.bias'], a1, pre_add=self.film[:, off:off + cout],
                     label=f'{n}.norm1')
        out = self._new(B, Ho, Ho, cout)
        if blk.skip_conv:
            skip_src = [xr] if resample else xs
            self._gemm([a1] + skip_src, f'{n}.conv1skip', cout, out, bias=W_[f'{n}.conv1skip.b'], out_scale=cfg.skip_scale,
                       label=f'{n}.conv1+skip')
        else:
            assert len(xs) == 1
            self._gemm([a1], f'{n}.conv1', cout, out, bias=W_[f'{n}.conv1.b'], residual=xr if resample else xs[0],
                       out_scale=cfg.skip_scale, label=f'{n}.conv1')
        if not blk.attention:
            return out
        heads, L = blk.heads, Ho * Ho
        if cout // heads != 64:
            raise NotImplementedError(f'{n}: the precise path implements head_dim-64 attention only')
        a2 = self._act('a1', B, Ho, Ho, cout)
        self._gn([out], cout, W_[f'{n}.norm2.weight'], W_[f'{n}.norm2.bias'], a2, silu=False, label=f'{n}.norm2')
        qkv = self._act('qkv', B, Ho, Ho, 3 * cout)
        self._gemm([a2], f'{n}.qkv', 3 * cout, qkv, bias=W_[f'{n}.qkv.b'], label=f'{n}.qkv')
        att = self._act('a0', B, Ho, Ho, cout)
        P.add_attention_prec(qkv.view(B * L, 6 * cout), att.view(B * L, 2 * cout), B, heads, L, cout, label=f'{n}.attn')
        out2 = self._new(B, Ho, Ho, cout)
        self._gemm([att], f'{n}.proj', cout, out2, bias=W_[f'{n}.proj.b'], residual=out, out_scale=cfg.skip_scale,
                   label=f'{n}.proj')
        return out2


class PreciseUNetEngine:
    """Split-fp16 twin of `UNetEngine`: shares the fp32 parameters (embedding MLP, norms, biases) of `base`, packs its
    own split conv / projection weights from the state dict."""

    def __init__(self, base: UNetEngine, state_dict: Dict[str, torch.Tensor], use_graphs: bool = True):
        self.base, self.cfg, self.device, self.use_graphs = base, base.cfg, base.device, use_graphs
        sd = {k[len('model.'):] if k.startswith('model.') else k: v for k, v in state_dict.items()}
        self.w: Dict[str, tuple] = {}
        if os.environ.get('B200NS_PREC_NOLO') == '1':
            from . import _lib
            _lib.check(_lib.lib().b200ns_debug_prec_nolo(1), 'debug_prec_nolo')
        self._pack(sd)
        self._plans: Dict[tuple, PreciseForwardPlan] = {}

    def _put(self, key: str, blocks, n_pad=None):
        w, acc_scale, segs = pack_split(blocks, n_pad)
        self.w[key] = (w.to(self.device), acc_scale, segs)

    def _pack(self, sd):
        cfg = self.cfg
        # channel split of every decoder block's concat input: recover it by replaying the skip stack (unet.py:_item)
        skips: List[int] = []
        c_cur = 0
        for b in cfg.enc + cfg.dec:
            n = b.name
            if b.kind == 'conv':
                wt = sd[f'{n}.weight'].detach().float().cpu()                      # [Cout, Cin, 3, 3], 9*Cin <= 64
                blk = torch.zeros(b.cout, 1, 64)
                blk[:, 0, :9 * b.cin] = wt.permute(0, 2, 3, 1).reshape(b.cout, -1)
                self._put(n, [blk])
                c_cur = b.cout
                skips.append(c_cur)
                continue
            if b.kind == 'aux_norm':
                continue
            if b.kind == 'aux_conv':
                self._put(n, [_conv_block(sd[f'{n}.weight'])], n_pad=16)
                continue
            srcs = [c_cur]
            if n.startswith('dec.') and c_cur != b.cin:
                srcs.append(skips.pop())
                assert sum(srcs) == b.cin, (n, srcs, b.cin)
            # conv0 reads the normalised concat materialised by gn_apply: one source
            self._put(f'{n}.conv0', [_conv_block(sd[f'{n}.conv0.weight'])])
            w1 = _conv_block(sd[f'{n}.conv1.weight'])
            if b.skip_conv:
                ws = sd[f'{n}.skip.weight'].detach().float().cpu()
                resampled = b.up or b.down                     # the skip then reads the single resampled tensor xr
                parts, c0 = [], 0
                for c in ([b.cin] if resampled else srcs):
                    parts.append(_conv_block(ws, c0, c0 + c))
                    c0 += c
                self._put(f'{n}.conv1skip', [w1] + parts)
            else:
                self._put(f'{n}.conv1', [w1])
            if b.attention:
                C = b.cout
                wq = sd[f'{n}.qkv.weight'].detach().float().cpu()[:, :, 0, 0].reshape(C, 3, C).permute(1, 0, 2)
                self._put(f'{n}.qkv', [wq.reshape(3 * C, 1, C).contiguous()])
                self._put(f'{n}.proj', [sd[f'{n}.proj.weight'].detach().float().cpu()[:, :, 0, 0].reshape(C, 1, C).contiguous()])
            c_cur = b.cout
            if n.startswith('enc.'):
                skips.append(c_cur)
        if cfg.model_type == 'DhariwalUNet':
            self._put('out_conv', [_conv_block(sd['out_conv.weight'])], n_pad=16)

    def plan(self, B: int, b_emb: int) -> PreciseForwardPlan:
        key = (B, b_emb)
        if key not in self._plans:
            self._plans[key] = PreciseForwardPlan(self, B, b_emb)
        return self._plans[key]

    def forward(self, x_in: torch.Tensor, c_noise: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                b_emb: Optional[int] = None) -> torch.Tensor:
        """x_in fp32 NCHW (already scaled by c_in) -> F_x fp32 NCHW view (tests; the search loop drives plans directly)."""
        B = x_in.shape[0]
        if b_emb is None:
            b_emb = class_labels.shape[0] if (class_labels is not None and class_labels.dim() == 2) else 1
        fp = self.plan(B, b_emb)
        fp.x_in.copy_(x_in)
        fp.emb_in.copy_(self.base.positional_embedding(c_noise.reshape(-1)).expand(fp.b_emb, -1))
        if self.cfg.label_dim:
            fp.labels.copy_(class_labels.to(torch.float32).reshape(-1, self.cfg.label_dim).expand(fp.b_emb, -1))
        fp.plan.run()
        return fp.out.permute(0, 3, 1, 2)

"""U-Net engine: the DhariwalUNet (ADM) / SongUNet (DDPM++) forward of the reference
(edm/training/networks.py:372-461, 229-363) as a static plan of sm_100a kernels.

The engine never imports or subclasses the reference classes: the network is described by its
`state_dict()` alone (key names + shapes), so it works for nets re-created from the pickles'
embedded source (edm/torch_utils/persistence.py) as well as for plain dicts of tensors.

Data flow per UNetBlock (networks.py:166-187), all activations bf16 NHWC:
  gn_stats(x[,skip]) -> gn_apply(norm0, SiLU, 2x resample)      -> a0   (+ xr: resampled raw x)
  tcgen05 conv3x3(a0) + bias                                    -> h
  gn_stats(h) -> gn_apply(norm1, FiLM scale/shift, SiLU)        -> a1
  tcgen05 [conv3x3(a1) | conv1x1(orig)] + bias (+orig) * skip_scale -> out   (one accumulator)
  [attention] gn_stats/apply(norm2) -> tcgen05 qkv ([Q|K|V] row-major) -> flash attention (V as MN-major operand)
              -> tcgen05 proj + bias + out, * skip_scale
The embedding MLP and all per-block `affine` layers depend only on (sigma, label), i.e. are
identical for all N candidates of an image: they run once per distinct image (b_emb rows).
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from ._lib import ACT_DTYPE

from .ops import Plan, pack_conv_up2


@dataclass
class Block:
    name: str
    kind: str            # 'conv' | 'block' | 'aux_norm' | 'aux_conv'
    cin: int
    cout: int
    res: int             # output resolution
    up: bool = False
    down: bool = False
    attention: bool = False
    heads: int = 0
    skip_conv: bool = False


@dataclass
class NetConfig:
    model_type: str      # 'DhariwalUNet' | 'SongUNet'
    img_resolution: int
    in_channels: int
    out_channels: int
    label_dim: int
    noise_channels: int
    emb_channels: int
    adaptive_scale: bool
    skip_scale: float
    eps: float
    enc: List[Block] = field(default_factory=list)
    dec: List[Block] = field(default_factory=list)


_KEY = re.compile(r'^(enc|dec)\.(\d+)x\d+_([a-z_]+\d*)\.')


def derive_config(sd: Dict[str, torch.Tensor]) -> NetConfig:
    """Recover the architecture from state-dict keys/shapes (module registration order ==
    execution order, networks.py:405-433 / 285-318)."""
    adm = 'out_conv.weight' in sd
    if not adm and not any('aux_conv' in k for k in sd):
        raise ValueError('state dict is neither a DhariwalUNet nor a SongUNet')
    order: List[str] = []
    for k in sd:
        m = _KEY.match(k)
        if m:
            prefix = k[:m.end() - 1]
            if prefix not in order:
                order.append(prefix)

    def exec_order(prefix: str):
        """Execution order (== module registration order, networks.py:405-433): encoder from the
        highest resolution down, decoder back up; do not rely on the dict's own ordering."""
        m = _KEY.match(prefix + '.')
        part, res, leaf = m.group(1), int(m.group(2)), m.group(3)
        stem = leaf.rstrip('0123456789')
        num = int(leaf[len(stem):]) if len(leaf) > len(stem) else 0
        rank = {'conv': 0, 'down': 0, 'in': 0, 'up': 1, 'block': 2, 'aux_up': 3, 'aux_down': 3, 'aux_skip': 3,
                'aux_residual': 3, 'aux_norm': 4, 'aux_conv': 5}[stem]
        return (0, -res, rank, num) if part == 'enc' else (1, res, rank, num)

    order.sort(key=exec_order)
    blocks: List[Block] = []
    for prefix in order:
        m = _KEY.match(prefix + '.')
        res, leaf = int(m.group(2)), m.group(3)
        if leaf == 'conv' or leaf == 'aux_conv':
            w = sd[f'{prefix}.weight']
            blocks.append(Block(prefix, 'conv' if leaf == 'conv' else 'aux_conv', w.shape[1], w.shape[0], res))
        elif leaf == 'aux_norm':
            c = sd[f'{prefix}.weight'].shape[0]
            blocks.append(Block(prefix, 'aux_norm', c, c, res))
        elif leaf.startswith(('aux_', )):
            raise NotImplementedError(f'{prefix}: NCSN++ skip/residual encoder-decoder variants are out of scope')
        else:
            w0 = sd[f'{prefix}.conv0.weight']
            cout, cin = w0.shape[0], w0.shape[1]
            attn = f'{prefix}.qkv.weight' in sd
            blocks.append(Block(prefix, 'block', cin, cout, res, up=leaf == 'up', down=leaf == 'down',
                                attention=attn, heads=(cout // 64 if adm else 1) if attn else 0,
                                skip_conv=f'{prefix}.skip.weight' in sd))
    first = blocks[0]
    aff = next(b for b in blocks if b.kind == 'block')
    cfg = NetConfig(model_type='DhariwalUNet' if adm else 'SongUNet', img_resolution=first.res,
                    in_channels=first.cin,
                    out_channels=(sd['out_conv.weight'].shape[0] if adm else
                                  sd[[b for b in blocks if b.kind == 'aux_conv'][-1].name + '.weight'].shape[0]),
                    label_dim=sd['map_label.weight'].shape[1] if 'map_label.weight' in sd else 0,
                    noise_channels=sd['map_layer0.weight'].shape[1], emb_channels=sd['map_layer0.weight'].shape[0],
                    adaptive_scale=sd[f'{aff.name}.affine.weight'].shape[0] == 2 * aff.cout,
                    skip_scale=1.0 if adm else math.sqrt(0.5), eps=1e-5 if adm else 1e-6)
    cfg.enc = [b for b in blocks if b.name.startswith('enc.')]
    cfg.dec = [b for b in blocks if b.name.startswith('dec.')]
    return cfg


def _pack_conv(w: torch.Tensor, splits: Optional[List[int]] = None) -> torch.Tensor:
    """[Cout,Cin,k,k] fp32 -> bf16 [Cout, K]; K = per input-channel split: (tap, channel)."""
    parts = []
    c0 = 0
    for c in (splits or [w.shape[1]]):
        ws = w[:, c0:c0 + c]
        parts.append(ws.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
        c0 += c
    return torch.cat(parts, dim=1).contiguous().to(ACT_DTYPE)


def _groups(c: int) -> int:
    return min(32, c // 4)                # networks.py:99


class ForwardPlan:
    """All buffers + the kernel plan for one (batch B, b_emb) shape.

    Lanes: at resolutions >= eng.lane_min_res the batch is processed as `eng.lanes` independent sub-batches whose ops
    sit on parallel branches of the captured CUDA graph, so the HBM-bound GroupNorm passes of one sub-batch overlap
    the tensor-core GEMMs of the other (and one sub-batch's GEMM fills the tail wave of the other's).  Every
    per-sample result is batch-invariant, so the split changes no bits.  Builder state (x, skips) always holds the
    FULL-batch tensors; ops see the current lane's slice (`_view`)."""

    # GroupNorm of the A operand inside the qkv GEMM (norm2 -> qkv, `Plan.add_gemm(a_norm=...)`): built, bit-identical,
    # and measured SLOWER than the separate pass at every level, so it is off unless B200NS_FUSED_NORM_A=1.
    # (class attribute: subclasses with their own __init__ -- classifier, SD, VAE plans -- inherit it)
    fused_norm_a = os.environ.get('B200NS_FUSED_NORM_A', '0') == '1'


    def __init__(self, eng: 'UNetEngine', B: int, b_emb: int):
        cfg, dev = eng.cfg, eng.device
        self.B, self.B_full, self.b_emb = B, B, b_emb
        self.fused_gn_stats = eng.fused_gn_stats
        H = cfg.img_resolution
        f32 = dict(device=dev, dtype=torch.float32)
        self.x_in = torch.zeros(B, cfg.in_channels, H, H, **f32)                  # c_in * x  (NCHW fp32)
        self.emb_in = torch.zeros(b_emb, cfg.noise_channels, **f32)               # positional embedding
        self.labels = torch.zeros(b_emb, max(cfg.label_dim, 1), **f32)
        self.out = torch.empty(B, H, H, cfg.out_channels, **f32)                  # F_x, NHWC fp32
        self.plan = Plan()
        self._scratch: Dict[str, torch.Tensor] = {}
        self._full: Dict[str, torch.Tensor] = {}          # persistent full-batch tensors by name
        self.block_out: Dict[str, torch.Tensor] = {}      # per-block outputs (persistent; per-layer parity tests)
        # (data_ptr, batch) -> fp32 [M/64, C, 2] per-channel (sum, sumsq) left behind by the GEMM that produced it
        self._stats: Dict[tuple, torch.Tensor] = {}
        self._dir: Dict[tuple, bool] = {}                 # (data_ptr, batch) -> tensor was written last-to-first
        self.alternate_walk = getattr(eng, 'alternate_walk', True)
        lanes = getattr(eng, 'lanes', 1)
        self.n_lanes = lanes if (lanes > 1 and B % lanes == 0 and (B // lanes) % b_emb == 0) else 1
        self.lane_min_res = getattr(eng, 'lane_min_res', 32)
        self._lane: Optional[int] = None                  # None: full batch on the main stream
        self._build(eng)
        if os.environ.get('B200NS_PDL') is None and self.B_full <= 4:
            self.plan.set_pdl(1)          # small batches are launch-latency bound: overlap each kernel's prologue with its predecessor
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            self.plan.instantiate_graph()

    # -- helpers
    def _buf(self, key: str, numel: int, dtype=ACT_DTYPE) -> torch.Tensor:
        key = f'{key}@{self._lane}'                       # scratch is private to a lane (lanes run concurrently)
        t = self._scratch.get(key)
        if t is None or t.numel() < numel:
            t = torch.empty(numel, device=self.x_in.device, dtype=dtype)
            self._scratch[key] = t
        return t

    def _act(self, key: str, B, H, W, C) -> torch.Tensor:
        return self._buf(key, B * H * W * C)[:B * H * W * C].view(B, H, W, C)

    def _persist(self, key: str, *shape, dtype=ACT_DTYPE) -> torch.Tensor:
        """Full-batch tensor that outlives the op (block outputs / skips), allocated once per name."""
        t = self._full.get(key)
        if t is None:
            t = torch.empty(self.B_full, *shape, device=self.x_in.device, dtype=dtype)
            self._full[key] = t
        return t

    def _view(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """The current lane's rows of a full-batch tensor."""
        if t is None or self._lane is None:
            return t
        Bs = self.B_full // self.n_lanes
        return t[self._lane * Bs:(self._lane + 1) * Bs]

    def _set_lane(self, lane: Optional[int]):
        self._lane = lane
        self.B = self.B_full if lane is None else self.B_full // self.n_lanes
        self.plan.set_lane(0 if lane is None else lane + 1)

    @staticmethod
    def _tk(t: torch.Tensor) -> tuple:
        return (t.data_ptr(), t.shape[0])

    def _rev(self, src: torch.Tensor, dst: Optional[torch.Tensor] = None, *more) -> bool:
        """Walk direction for an op reading `src`: opposite to the direction `src` was written in, so the op starts on
        the rows its producer finished last (still in L2).  Records the direction on the op's outputs."""
        rev = self.alternate_walk and not self._dir.get(self._tk(src), False)
        for t in (dst,) + more:
            if t is not None:
                self._dir[self._tk(t)] = rev
        return rev

    def _new_stats(self, t: torch.Tensor, key: Optional[str] = None, full: Optional[torch.Tensor] = None):
        """Statistics buffer for a GEMM output `t` that a GroupNorm will consume (None when the fused path is off).
        `key`: lane-private scratch name; `full`: the full-batch tensor `t` is a lane view of (the statistics of a
        persistent tensor are persistent and full-batch too, so a full-batch consumer can read all lanes' rows)."""
        if not self.fused_gn_stats:
            return None
        B, H, W, C = t.shape
        rows = B * H * W // 64
        if key is not None:
            st = self._buf(key, rows * C * 2, torch.float32)[:rows * C * 2]
        else:
            full = t if full is None else full
            fk = ('stats',) + self._tk(full)
            sf = self._full.get(fk)
            if sf is None:
                sf = torch.empty(full.shape[0] * H * W // 64, C, 2, device=t.device, dtype=torch.float32)
                self._full[fk] = sf
                self._stats[self._tk(full)] = sf
            st = sf if self._lane is None else sf[self._lane * rows:(self._lane + 1) * rows]
        self._stats[self._tk(t)] = st
        return st

    @staticmethod
    def _num_groups(C: int) -> int:
        return _groups(C)                 # EDM: min(32, C // 4) (networks.py:99)

    @staticmethod
    def _splits(HW: int) -> int:
        """GroupNorm partial-sum splits: a function of the image size ONLY, so the reduction order -- and
        with it every bit of the network output -- is independent of batch size and batch position."""
        return max(1, min(16, HW // 64))

    def _gn(self, xs, C, H, W, gamma, beta, out, *, silu=True, resample=0, raw_out=None, film=None, pre_add=None,
            label='', eps=None):
        g = self._num_groups(C)
        eps = self._eps if eps is None else eps
        stats = [self._stats.get(self._tk(t)) for t in xs]
        if all(st is not None for st in stats):
            # statistics come from the producing GEMMs' epilogues: no pass over the activations
            mr = self._buf('mean_rstd', self.B * 64 * 2, torch.float32)[:self.B * g * 2]
            self.plan.add_gn_norm(stats, xs, g, eps, mr, gamma, beta, out, pre_add=pre_add,
                                  film_scale=film[0] if film else None, film_shift=film[1] if film else None,
                                  b_emb=self.b_emb, silu=silu, resample=resample, raw_out=raw_out,
                                  reverse=self._rev(xs[0], out, raw_out), label=label)
            return
        splits = self._splits(H * W)
        partial = self._buf('partial', self.B * 512 * 32 * 2, torch.float64)[:self.B * splits * g * 2].view(
            self.B, splits, g, 2)
        self.plan.add_gn_stats(xs, g, partial, splits, pre_add=pre_add, b_emb=self.b_emb, label=f'{label}.stats')
        self.plan.add_gn_apply(xs, g, partial, splits, eps, gamma, beta, out, pre_add=pre_add,
                               film_scale=film[0] if film else None, film_shift=film[1] if film else None,
                               b_emb=self.b_emb, silu=silu, resample=resample, raw_out=raw_out,
                               reverse=self._rev(xs[0], out, raw_out), label=f'{label}.apply')

    # -- build
    def _build(self, eng: 'UNetEngine'):
        cfg, P, W_ = eng.cfg, self.plan, eng.w
        b_emb = self.b_emb
        adm = cfg.model_type == 'DhariwalUNet'
        self._eps = cfg.eps
        dev = self.x_in.device
        f32 = dict(device=dev, dtype=torch.float32)
        E = cfg.emb_channels

        # ---- embedding network (candidate-invariant: b_emb rows), main stream
        t0 = torch.empty(b_emb, E, **f32)
        self.emb = torch.empty(b_emb, E, **f32)
        if adm:
            P.add_linear(self.emb_in, W_['map_layer0.weight'], t0, bias=W_['map_layer0.bias'], act=1, label='map_layer0')
            lab = None
            if cfg.label_dim:
                lab = torch.empty(b_emb, E, **f32)
                P.add_linear(self.labels, W_['map_label.weight'], lab, label='map_label')
            P.add_linear(t0, W_['map_layer1.weight'], self.emb, bias=W_['map_layer1.bias'], add=lab, act=1,
                         label='map_layer1')
        else:
            src = self.emb_in
            if cfg.label_dim:
                src = torch.empty(b_emb, cfg.noise_channels, **f32)
                P.add_linear(self.labels, W_['map_label.weight_scaled'], src, bias=W_['map_label.bias'],
                             add=self.emb_in, label='map_label')
            P.add_linear(src, W_['map_layer0.weight'], t0, bias=W_['map_layer0.bias'], act=1, label='map_layer0')
            P.add_linear(t0, W_['map_layer1.weight'], self.emb, bias=W_['map_layer1.bias'], act=1, label='map_layer1')
        self.film = torch.empty(b_emb, eng.affine_total, **f32)
        P.add_linear(self.emb, W_['affine_all.weight'], self.film, bias=W_['affine_all.bias'], label='affine_all')

        # ---- U-Net body as regions of consecutive items that are either split over the lanes or run full-batch
        items = [('enc', blk) for blk in cfg.enc] + [('dec', blk) for blk in cfg.dec] + ([('out', None)] if adm else [])
        res_of = lambda it: cfg.img_resolution if it[1] is None else it[1].res
        regions: List[tuple] = []
        for it in items:
            split = self.n_lanes > 1 and res_of(it) >= self.lane_min_res
            if regions and regions[-1][0] == split:
                regions[-1][1].append(it)
            else:
                regions.append((split, [it]))
        state = dict(x=None, skips=[], aux_in=None)
        for split, its in regions:
            start = dict(x=state['x'], skips=list(state['skips']), aux_in=None)
            for lane in (range(self.n_lanes) if split else [None]):
                self._set_lane(lane)
                state = dict(x=start['x'], skips=list(start['skips']), aux_in=None)     # every lane replays the region
                for kind, blk in its:
                    self._item(eng, kind, blk, state)
        self._set_lane(None)

    def _item(self, eng: 'UNetEngine', kind: str, blk: Optional[Block], st: dict):
        cfg, P, W_ = eng.cfg, self.plan, eng.w
        B = self.B
        if kind == 'out':                                    # ADM: out_norm + SiLU + out_conv (networks.py:460)
            H = cfg.img_resolution
            x = self._view(st['x'])
            C = x.shape[3]
            a = self._act('a0', B, H, H, C)
            self._gn([x], C, H, H, W_['out_norm.weight'], W_['out_norm.bias'], a, silu=True, label='out_norm')
            P.add_gemm([a], [(0, 9, 0, C // 64)], W_['out_conv.w'], cfg.out_channels, self._view(self.out),
                       bias=W_['out_conv.b'], reverse=self._rev(a), label='out_conv')
        elif kind == 'enc':
            if blk.kind == 'conv':
                H = blk.res
                col = self._act('col', B, H, H, 64)
                P.add_im2col(self._view(self.x_in), col, label=f'{blk.name}.im2col')
                xf = self._persist(blk.name, H, H, blk.cout)
                x = self._view(xf)
                P.add_gemm([col], [(0, 1, 0, 1)], W_[f'{blk.name}.w'], blk.cout, x, bias=W_[f'{blk.name}.b'],
                           alg_k=9 * blk.cin, gn_stats=self._new_stats(x, full=xf), reverse=self._rev(col, x),
                           label=f'{blk.name}')
                st['x'] = xf
            else:
                st['x'] = self._block(eng, blk, [st['x']])
            self.block_out[blk.name] = st['x']
            st['skips'].append(st['x'])
        elif blk.kind == 'aux_norm':
            H = blk.res
            st['aux_in'] = self._act('a0', B, H, H, blk.cin)
            self._gn([self._view(st['x'])], blk.cin, H, H, W_[f'{blk.name}.weight'], W_[f'{blk.name}.bias'], st['aux_in'],
                     silu=True, label=blk.name)
        elif blk.kind == 'aux_conv':
            P.add_gemm([st['aux_in']], [(0, 9, 0, blk.cin // 64)], W_[f'{blk.name}.w'], blk.cout, self._view(self.out),
                       bias=W_[f'{blk.name}.b'], reverse=self._rev(st['aux_in']), label=blk.name)
        else:
            xs = [st['x']]
            if st['x'].shape[3] != blk.cin:
                xs.append(st['skips'].pop())
                assert xs[0].shape[3] + xs[1].shape[3] == blk.cin
            st['x'] = self._block(eng, blk, xs)
            self.block_out[blk.name] = st['x']

    def _block(self, eng: 'UNetEngine', blk: Block, xs: List[torch.Tensor]) -> torch.Tensor:
        cfg, P, W_ = eng.cfg, self.plan, eng.w
        B, n = self.B, blk.name
        xs = [self._view(t) for t in xs]                  # builder state holds full-batch tensors
        Hin = xs[0].shape[1]
        Ho = blk.res
        cin, cout = blk.cin, blk.cout
        resample = 1 if blk.up else (2 if blk.down else 0)
        need_raw = resample != 0                      # skip path sees the resampled raw input
        xr = self._act('xr', B, Ho, Ho, cin) if need_raw else None
        h = self._act('h', B, Ho, Ho, cout)
        if blk.up and eng.fused_upsample and len(xs) == 1:
            # Conv2d(up=True) with resample_filter [1,1] (networks.py:72-80) = conv3x3(nearest_up2(.)): normalise at the LOW
            # resolution and let four 2x2-tap phase GEMMs write the high-res conv0 output (16 instead of 36 MACs per input
            # channel and output pixel; the upsampled activation is never written).  The skip path still sees up2(x).
            a0 = self._act('a0', B, Hin, Hin, cin)
            self._gn(xs, cin, Hin, Hin, W_[f'{n}.norm0.weight'], W_[f'{n}.norm0.bias'], a0, silu=True, label=f'{n}.norm0')
            P.add_upsample2x(xs[0], xr, label=f'{n}.skip_up2')
            P.add_gemm([a0], [(0, 9, 0, cin // 64)], W_[f'{n}.conv0.wup'], cout, h, bias=W_[f'{n}.conv0.b'],
                       gn_stats=self._new_stats(h, 'h_stats'), reverse=self._rev(a0, h), label=f'{n}.conv0', upsample2x=True)
        else:
            a0 = self._act('a0', B, Ho, Ho, cin)
            self._gn(xs, cin, Hin, Hin, W_[f'{n}.norm0.weight'], W_[f'{n}.norm0.bias'], a0, silu=True, resample=resample,
                     raw_out=xr, label=f'{n}.norm0')
            P.add_gemm([a0], [(0, 9, 0, cin // 64)], W_[f'{n}.conv0.w'], cout, h, bias=W_[f'{n}.conv0.b'],
                       gn_stats=self._new_stats(h, 'h_stats'), reverse=self._rev(a0, h), label=f'{n}.conv0')
        a1 = self._act('a1', B, Ho, Ho, cout)
        off = eng.affine_off[n]
        if cfg.adaptive_scale:
            film = (self.film[:, off:off + cout], self.film[:, off + cout:off + 2 * cout])
            self._gn([h], cout, Ho, Ho, W_[f'{n}.norm1.weight'], W_[f'{n}.norm1.bias'], a1, film=film, label=f'{n}.norm1')
        else:
            self._gn([h], cout, Ho, Ho, W_[f'{n}.norm1.weight'], W_[f'{n}.norm1.bias'], a1,
                     pre_add=self.film[:, off:off + cout], label=f'{n}.norm1')
        out_full = self._persist(f'{n}.out', Ho, Ho, cout)
        out = self._view(out_full)
        if blk.skip_conv:
            skip_src = [xr] if need_raw else xs
            srcs = [a1] + skip_src
            segs = [(0, 9, 0, cout // 64)] + [(i + 1, 1, 0, t.shape[3] // 64) for i, t in enumerate(skip_src)]
            P.add_gemm(srcs, segs, W_[f'{n}.conv1skip.w'], cout, out, bias=W_[f'{n}.conv1skip.b'],
                       out_scale=cfg.skip_scale, gn_stats=self._new_stats(out, full=out_full), reverse=self._rev(a1, out),
                       label=f'{n}.conv1+skip')
        else:
            res = xr if need_raw else xs[0]
            assert len(xs) == 1
            P.add_gemm([a1], [(0, 9, 0, cout // 64)], W_[f'{n}.conv1.w'], cout, out, bias=W_[f'{n}.conv1.b'],
                       residual=res, out_scale=cfg.skip_scale, gn_stats=self._new_stats(out, full=out_full),
                       reverse=self._rev(a1, out),
                       label=f'{n}.conv1')
        if not blk.attention:
            return out_full
        heads, L = blk.heads, Ho * Ho
        hd = cout // heads
        if hd not in (64, 256) or (hd == 256 and (heads != 1 or L > 256)):
            raise NotImplementedError(f'{n}: attention head_dim {hd} x {heads} heads at L={L} is not implemented')
        qkv = self._act('qkv', B, Ho, Ho, 3 * cout)          # [Q | K | V], each head-major, row-major per pixel
        st2 = self._stats.get(self._tk(out))
        L2 = Ho * Ho
        if self.fused_norm_a and st2 is not None and (L2 == 64 or L2 % 128 == 0):
            # qkv(norm2(x)) (networks.py:182-183) with the GroupNorm applied INSIDE the GEMM's operand path: four transform
            # warps normalise each TMA-landed A stage in shared memory before the MMA warp consumes it, so the normalised
            # tensor is never written or re-read (bit-identical to the separate gn_apply pass)
            g = self._num_groups(cout)
            mr = self._buf('mean_rstd', self.B * 64 * 2, torch.float32)[:self.B * g * 2]
            P.add_gn_finalize([st2], [cout], self.B, L2, g, self._eps, mr, b_emb=self.b_emb, label=f'{n}.norm2.finalize')
            P.add_gemm([out], [(0, 1, 0, cout // 64)], W_[f'{n}.qkv.w'], 3 * cout, qkv, bias=W_[f'{n}.qkv.b'],
                       reverse=self._rev(out, qkv), label=f'{n}.norm2+qkv',
                       a_norm=(mr, W_[f'{n}.norm2.weight'], W_[f'{n}.norm2.bias'], g))
        else:
            a2 = self._act('a1', B, Ho, Ho, cout)
            self._gn([out], cout, Ho, Ho, W_[f'{n}.norm2.weight'], W_[f'{n}.norm2.bias'], a2, silu=False, label=f'{n}.norm2')
            P.add_gemm([a2], [(0, 1, 0, cout // 64)], W_[f'{n}.qkv.w'], 3 * cout, qkv, bias=W_[f'{n}.qkv.b'],
                       reverse=self._rev(a2, qkv), label=f'{n}.qkv')
        att = self._act('a0', B, Ho, Ho, cout)
        # V is consumed in place as an MN-major UMMA operand: no transposed copy
        P.add_attention(qkv.view(B * L, 3 * cout), cout, None, att.view(B * L, cout), B, heads, L,
                        v_col0=2 * cout, head_dim=hd, reverse=self._rev(qkv, att), label=f'{n}.attn')
        out2_full = self._persist(f'{n}.out2', Ho, Ho, cout)
        out2 = self._view(out2_full)
        P.add_gemm([att], [(0, 1, 0, cout // 64)], W_[f'{n}.proj.w'], cout, out2, bias=W_[f'{n}.proj.b'], residual=out,
                   out_scale=cfg.skip_scale, gn_stats=self._new_stats(out2, full=out2_full), reverse=self._rev(att, out2),
                   label=f'{n}.proj')
        return out2_full


class UNetEngine:
    """Packed weights + cached ForwardPlans.  `forward(x_in, c_noise, labels)` returns F_x."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', use_graphs: bool = True,
                 fused_gn_stats: bool = True, alternate_walk: bool = True, lanes: Optional[int] = None,
                 lane_min_res: int = 32, fused_upsample: Optional[bool] = None):
        from . import _lib
        self.use_graphs = use_graphs
        # up-sampling blocks: conv0(up2(x)) as four 2x2-tap phase GEMMs over the low-res input (B200NS_FUSED_UP=0: off)
        self.fused_upsample = (os.environ.get('B200NS_FUSED_UP', '1') != '0') if fused_upsample is None else fused_upsample
        self.fused_gn_stats = fused_gn_stats      # GroupNorm statistics from the producing GEMM's epilogue
        self.alternate_walk = alternate_walk      # consecutive kernels walk the batch in opposite directions (L2 reuse)
        # sub-batches on parallel graph branches at resolutions >= lane_min_res (see ForwardPlan)
        self.lanes = int(os.environ.get('B200NS_LANES', '1')) if lanes is None else lanes
        self.lane_min_res = lane_min_res
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('UNetEngine requires a CUDA device (B200); there is no CPU fallback')
        _lib.lib()
        sd = {k[len('model.'):] if k.startswith('model.') else k: v for k, v in state_dict.items()}
        self.cfg = derive_config(sd)
        for b in self.cfg.enc + self.cfg.dec:
            if b.kind == 'block' and (b.cin % 64 or b.cout % 64):
                raise NotImplementedError(f'{b.name}: channel counts must be multiples of 64 (got {b.cin}->{b.cout})')
        self.w: Dict[str, torch.Tensor] = {}
        self.affine_off: Dict[str, int] = {}
        self._pack(sd)
        self._plans: Dict[tuple, ForwardPlan] = {}

    # -- weights
    def _pack(self, sd):
        cfg, dev, w = self.cfg, self.device, self.w
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        for k in ('map_layer0.weight', 'map_layer0.bias', 'map_layer1.weight', 'map_layer1.bias'):
            w[k] = f(sd[k])
        if cfg.label_dim:
            if cfg.model_type == 'DhariwalUNet':
                w['map_label.weight'] = f(sd['map_label.weight'])
            else:       # networks.py:327: map_label(labels * sqrt(in_features))
                w['map_label.weight_scaled'] = f(sd['map_label.weight']) * math.sqrt(cfg.label_dim)
                w['map_label.bias'] = f(sd['map_label.bias'])
        aff_w, aff_b, off = [], [], 0
        for b in cfg.enc + cfg.dec:
            n = b.name
            if b.kind == 'conv':
                wp = torch.zeros(b.cout, 64, dtype=ACT_DTYPE)
                wp[:, :9 * b.cin] = _pack_conv(sd[f'{n}.weight'].detach().float().cpu())
                w[f'{n}.w'], w[f'{n}.b'] = wp.to(dev), f(sd[f'{n}.bias'])
            elif b.kind == 'aux_norm':
                w[f'{n}.weight'], w[f'{n}.bias'] = f(sd[f'{n}.weight']), f(sd[f'{n}.bias'])
            elif b.kind == 'aux_conv':
                self._pack_out_conv(sd, n, n)
            else:
                for nm in ('norm0', 'norm1') + (('norm2',) if b.attention else ()):
                    w[f'{n}.{nm}.weight'], w[f'{n}.{nm}.bias'] = f(sd[f'{n}.{nm}.weight']), f(sd[f'{n}.{nm}.bias'])
                # conv0 reads the normalised concat materialised by gn_apply: plain (tap, channel) order
                if b.up and self.fused_upsample:
                    w[f'{n}.conv0.wup'] = pack_conv_up2(sd[f'{n}.conv0.weight']).to(dev)
                else:
                    w[f'{n}.conv0.w'] = _pack_conv(sd[f'{n}.conv0.weight'].detach().float().cpu()).to(dev)
                w[f'{n}.conv0.b'] = f(sd[f'{n}.conv0.bias'])
                w1 = _pack_conv(sd[f'{n}.conv1.weight'].detach().float().cpu())
                if b.skip_conv:
                    ws = sd[f'{n}.skip.weight'].detach().float().cpu()[:, :, 0, 0].to(ACT_DTYPE)
                    w[f'{n}.conv1skip.w'] = torch.cat([w1, ws], dim=1).contiguous().to(dev)
                    w[f'{n}.conv1skip.b'] = f(sd[f'{n}.conv1.bias']) + f(sd[f'{n}.skip.bias'])
                else:
                    w[f'{n}.conv1.w'], w[f'{n}.conv1.b'] = w1.to(dev), f(sd[f'{n}.conv1.bias'])
                if b.attention:
                    C = b.cout
                    # reference channel order (head, d, {q,k,v}) (networks.py:182) -> [Q | K | V], head-major
                    wq = sd[f'{n}.qkv.weight'].detach().float().cpu()[:, :, 0, 0].reshape(C, 3, C).permute(1, 0, 2)
                    bq = sd[f'{n}.qkv.bias'].detach().float().cpu().reshape(C, 3).permute(1, 0)
                    w[f'{n}.qkv.w'] = wq.reshape(3 * C, C).contiguous().to(ACT_DTYPE).to(dev)
                    w[f'{n}.qkv.b'] = bq.reshape(3 * C).contiguous().to(dev)
                    w[f'{n}.proj.w'] = sd[f'{n}.proj.weight'].detach().float().cpu()[:, :, 0, 0].contiguous().to(
                        ACT_DTYPE).to(dev)
                    w[f'{n}.proj.b'] = f(sd[f'{n}.proj.bias'])
                self.affine_off[n] = off
                aff_w.append(sd[f'{n}.affine.weight'].detach().float().cpu())
                aff_b.append(sd[f'{n}.affine.bias'].detach().float().cpu())
                off += aff_w[-1].shape[0]
        self.affine_total = off
        w['affine_all.weight'] = torch.cat(aff_w, dim=0).contiguous().to(dev)
        w['affine_all.bias'] = torch.cat(aff_b, dim=0).contiguous().to(dev)
        if cfg.model_type == 'DhariwalUNet':
            w['out_norm.weight'], w['out_norm.bias'] = f(sd['out_norm.weight']), f(sd['out_norm.bias'])
            self._pack_out_conv(sd, 'out_conv', 'out_conv')

    def _pack_out_conv(self, sd, key, name):
        wt = sd[f'{key}.weight'].detach().float().cpu()
        wp = torch.zeros(16, 9 * wt.shape[1], dtype=ACT_DTYPE)
        wp[:wt.shape[0]] = _pack_conv(wt)
        self.w[f'{name}.w'] = wp.to(self.device)
        self.w[f'{name}.b'] = sd[f'{key}.bias'].detach().to(device=self.device, dtype=torch.float32).contiguous()

    # -- forward
    def plan(self, B: int, b_emb: int) -> ForwardPlan:
        key = (B, b_emb)
        if key not in self._plans:
            self._plans[key] = ForwardPlan(self, B, b_emb)
        return self._plans[key]

    def positional_embedding(self, c_noise: torch.Tensor) -> torch.Tensor:
        """networks.py:200-206 (+ the sin/cos swap of SongUNet.forward :323).  [b_emb] -> [b_emb, C]."""
        cfg = self.cfg
        half = cfg.noise_channels // 2
        endpoint = cfg.model_type == 'SongUNet'
        freqs = torch.arange(0, half, dtype=torch.float32, device=c_noise.device)
        freqs = freqs / (half - (1 if endpoint else 0))
        freqs = (1 / 10000) ** freqs
        ang = torch.outer(c_noise.to(torch.float32), freqs)
        emb = torch.cat([ang.cos(), ang.sin()], dim=1)
        if endpoint:
            emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)
        return emb

    def run(self, fp: ForwardPlan, c_noise: torch.Tensor, class_labels: Optional[torch.Tensor]) -> torch.Tensor:
        """Run the plan on fp.x_in (already filled by the caller).  Returns fp.out (NHWC fp32)."""
        fp.emb_in.copy_(self.positional_embedding(c_noise.reshape(-1)).expand(fp.b_emb, -1))
        if self.cfg.label_dim:
            fp.labels.copy_(class_labels.to(torch.float32).reshape(-1, self.cfg.label_dim).expand(fp.b_emb, -1))
        fp.plan.run()
        return fp.out

    def forward(self, x_in: torch.Tensor, c_noise: torch.Tensor, class_labels: Optional[torch.Tensor] = None,
                b_emb: Optional[int] = None) -> torch.Tensor:
        """x_in fp32 NCHW [B,C,H,W] (already scaled by c_in) -> F_x fp32 NCHW (a view of the NHWC result)."""
        B = x_in.shape[0]
        if b_emb is None:
            b_emb = class_labels.shape[0] if (class_labels is not None and class_labels.dim() == 2) else 1
        if B % b_emb:
            raise ValueError('batch must be a multiple of the number of distinct embeddings')
        fp = self.plan(B, b_emb)
        fp.x_in.copy_(x_in)
        return self.run(fp, c_noise, class_labels).permute(0, 3, 1, 2)

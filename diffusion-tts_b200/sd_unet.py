"""SD-1.5-shaped `UNet2DConditionModel` (BASELINE.json config 5) as a static plan of sm_100a kernels.

Replaces, for the beam-search step of the SD backend, the vendored diffusers forward
(sd/diffusers/src/diffusers/models/unets/unet_2d_condition.py:1039-1310 and the blocks it calls: ResnetBlock2D,
Transformer2DModel / BasicTransformerBlock / AttnProcessor2_0, GEGLU, Downsample2D, Upsample2D).  Like the EDM engine
it is described by the state dict alone; activations are bf16 NHWC (a token tensor [B, H*W, C] IS the NHWC tensor).

Per ResnetBlock2D:   gn_finalize/apply(norm1, SiLU) -> conv1 (+ conv bias + time_emb_proj(silu(emb)) folded into the bias)
                     -> gn(norm2, SiLU) -> [conv2 | conv_shortcut(x)] in one accumulator (+ x when there is no shortcut)
Per Transformer2D:   gn(norm, eps 1e-6) -> proj_in -> LN1 -> fused qkv GEMM -> attention -> to_out (+residual)
                     -> LN2 -> q GEMM -> cross-attention over the 77 context tokens -> to_out (+residual)
                     -> LN3 -> GEGLU (GEMM to 8C, hidden*gelu(gate)) -> GEMM 4C->C (+residual) -> proj_out (+block input)
Heads: SD-1.5 uses 8 heads of C/8 = 40 / 80 / 160 channels.  The projection weights are packed with every head zero-padded
to Dp = 64 / 128 / 192, so Q K^T and P V are unchanged and the tcgen05 attention kernel sees whole swizzle atoms; the
softmax scale stays the TRUE head_dim^-0.5.
Candidate-invariant work runs once: the timestep embedding MLP and all 22 time_emb_proj layers (same t for every
candidate) are one fp32 linear per call; the cross-attention K/V projections of the two contexts ([uncond, cond]) are
computed once per prompt (`set_context`) and shared by all candidates, beams and steps.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from ._lib import ACT_DTYPE

from .ops import Plan, interleave_geglu, pack_conv_up2
from .unet import ForwardPlan, _pack_conv

CTX_ROWS = 128          # 77 context tokens padded to a multiple of the key tile


def sd_config_from_state_dict(sd: Dict[str, torch.Tensor]) -> dict:
    n_down = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('down_blocks.'))
    boc = [sd[f'down_blocks.{i}.resnets.0.conv1.weight'].shape[0] for i in range(n_down)]
    lpb = 1 + max(int(k.split('.')[3]) for k in sd if k.startswith('down_blocks.0.resnets.'))
    cross = [f'down_blocks.{i}.attentions.0.norm.weight' in sd for i in range(n_down)]
    cdim = next(v.shape[1] for k, v in sd.items() if k.endswith('attn2.to_k.weight'))
    return dict(block_out_channels=boc, layers_per_block=lpb, cross_attn_down=cross, cross_attention_dim=cdim, heads=8,
                in_channels=sd['conv_in.weight'].shape[1], out_channels=sd['conv_out.weight'].shape[0])


def _pad_dim(hd: int) -> int:
    dp = 64 * ((hd + 63) // 64)
    if dp > 192:
        raise NotImplementedError(f'attention head_dim {hd} > 192 is not supported by the SD engine')
    return dp


class SDPlan(ForwardPlan):
    """Buffers + kernel plan of the SD UNet for a fixed batch B (= 2 x candidates: [uncond half; cond half])."""

    def __init__(self, eng: 'SDUNetEngine', B: int, H: int):
        dev = eng.device
        cfg = eng.cfg
        self.B, self.B_full, self.b_emb = B, B, 1
        self.H = H
        f32 = dict(device=dev, dtype=torch.float32)
        self.x_in = torch.zeros(B, cfg['in_channels'], H, H, **f32)                 # latents, fp32 NCHW
        self.emb_in = torch.zeros(1, cfg['block_out_channels'][0], **f32)           # sinusoidal timestep embedding
        self.out = torch.empty(B, H, H, cfg['out_channels'], **f32)                 # eps, fp32 NHWC
        self.plan = Plan()
        self._scratch: Dict[str, torch.Tensor] = {}
        self._full: Dict[str, torch.Tensor] = {}
        self.block_out: Dict[str, torch.Tensor] = {}
        self._stats: Dict[tuple, torch.Tensor] = {}
        self._dir: Dict[tuple, bool] = {}
        self.fused_gn_stats = True
        self.alternate_walk = True
        self.n_lanes, self._lane = 1, None
        self._eps = 1e-5
        self._build_sd(eng)
        if os.environ.get('B200NS_PDL') is None and self.B_full <= 4:
            self.plan.set_pdl(1)          # small batches are launch-latency bound: overlap each kernel's prologue with its predecessor
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            self.plan.instantiate_graph()

    @staticmethod
    def _num_groups(C: int) -> int:
        return 32                         # norm_num_groups (unet_2d_condition.py:195)

    # ------------------------------------------------------------------ blocks
    def _resnet(self, eng, p: str, xs: List[torch.Tensor]) -> torch.Tensor:
        P, W_ = self.plan, eng.w
        B = self.B
        H = xs[0].shape[1]
        cin = sum(t.shape[3] for t in xs)
        cout = eng.cout[p]
        a0 = self._act('a0', B, H, H, cin)
        self._gn(xs, cin, H, H, W_[f'{p}.norm1.weight'], W_[f'{p}.norm1.bias'], a0, silu=True, label=f'{p}.norm1')
        h = self._act('h', B, H, H, cout)
        off = eng.temb_off[p]
        P.add_gemm([a0], [(0, 9, 0, cin // 64)], W_[f'{p}.conv1.w'], cout, h, bias=self.temb_bias[0, off:off + cout],
                   gn_stats=self._new_stats(h, 'h_stats'), reverse=self._rev(a0, h), label=f'{p}.conv1')
        a1 = self._act('a1', B, H, H, cout)
        self._gn([h], cout, H, H, W_[f'{p}.norm2.weight'], W_[f'{p}.norm2.bias'], a1, silu=True, label=f'{p}.norm2')
        out = self._persist(f'{p}.out', H, H, cout)
        if f'{p}.conv2sc.w' in W_:
            srcs = [a1] + xs
            segs = [(0, 9, 0, cout // 64)] + [(i + 1, 1, 0, t.shape[3] // 64) for i, t in enumerate(xs)]
            P.add_gemm(srcs, segs, W_[f'{p}.conv2sc.w'], cout, out, bias=W_[f'{p}.conv2sc.b'],
                       gn_stats=self._new_stats(out), reverse=self._rev(a1, out), label=f'{p}.conv2+shortcut')
        else:
            assert len(xs) == 1
            P.add_gemm([a1], [(0, 9, 0, cout // 64)], W_[f'{p}.conv2.w'], cout, out, bias=W_[f'{p}.conv2.b'], residual=xs[0],
                       gn_stats=self._new_stats(out), reverse=self._rev(a1, out), label=f'{p}.conv2')
        self.block_out[p] = out
        return out

    def _attn(self, eng, p: str, n: torch.Tensor, res: torch.Tensor, out: torch.Tensor, cross: bool):
        """attn1 (self) / attn2 (cross): n = LayerNorm'd tokens [B,H,W,C]; out = to_out(attention) + res."""
        P, W_ = self.plan, eng.w
        B, H, _, C = n.shape
        heads, L = eng.cfg['heads'], H * H
        hd = C // heads
        dp = _pad_dim(hd)
        scale = 1.0 / math.sqrt(hd)
        att = self._act('att', B, H, H, heads * dp)
        if not cross:
            qkv = self._act('qkv', B, H, H, 3 * heads * dp)
            P.add_gemm([n], [(0, 1, 0, C // 64)], W_[f'{p}.qkv.w'], 3 * heads * dp, qkv, reverse=self._rev(n, qkv),
                       label=f'{p}.qkv')
            P.add_attention(qkv.view(B * L, 3 * heads * dp), heads * dp, None, att.view(B * L, heads * dp), B, heads, L,
                            v_col0=2 * heads * dp, head_dim=dp, scale=scale, reverse=self._rev(qkv, att), label=f'{p}.attn')
        else:
            q = self._act('qkv', B, H, H, heads * dp)
            P.add_gemm([n], [(0, 1, 0, C // 64)], W_[f'{p}.q.w'], heads * dp, q, reverse=self._rev(n, q), label=f'{p}.q')
            kv = eng.ctx_kv[p]                                   # [2*CTX_ROWS, 2*heads*dp] = [K | V], set_context()
            P.add_attention(q.view(B * L, heads * dp), 0, None, att.view(B * L, heads * dp), B, heads, L, v_col0=heads * dp,
                            head_dim=dp, scale=scale, kv=kv, kv_rows=CTX_ROWS, kv_len=eng.ctx_len, kv_div=B // 2,
                            reverse=self._rev(q, att), label=f'{p}.xattn')
        P.add_gemm([att], [(0, 1, 0, heads * dp // 64)], W_[f'{p}.out.w'], C, out, bias=W_[f'{p}.out.b'], residual=res,
                   reverse=self._rev(att, out), label=f'{p}.to_out')

    def _transformer(self, eng, p: str, x: torch.Tensor) -> torch.Tensor:
        P, W_ = self.plan, eng.w
        B, H, _, C = x.shape
        a0 = self._act('a0', B, H, H, C)
        self._gn([x], C, H, H, W_[f'{p}.norm.weight'], W_[f'{p}.norm.bias'], a0, silu=False, eps=1e-6, label=f'{p}.norm')
        t0 = self._act('t0', B, H, H, C)
        P.add_gemm([a0], [(0, 1, 0, C // 64)], W_[f'{p}.proj_in.w'], C, t0, bias=W_[f'{p}.proj_in.b'],
                   reverse=self._rev(a0, t0), label=f'{p}.proj_in')
        t = f'{p}.transformer_blocks.0'
        n = self._act('ln', B, H, H, C)
        P.add_layernorm(t0, W_[f'{t}.norm1.weight'], W_[f'{t}.norm1.bias'], n, label=f'{t}.norm1')
        t1 = self._act('t1', B, H, H, C)
        self._attn(eng, f'{t}.attn1', n, t0, t1, cross=False)
        P.add_layernorm(t1, W_[f'{t}.norm2.weight'], W_[f'{t}.norm2.bias'], n, label=f'{t}.norm2')
        t2 = self._act('t0', B, H, H, C)
        self._attn(eng, f'{t}.attn2', n, t1, t2, cross=True)
        P.add_layernorm(t2, W_[f'{t}.norm3.weight'], W_[f'{t}.norm3.bias'], n, label=f'{t}.norm3')
        f = self._act('fff', B, H, H, 4 * C)
        if eng.fused_geglu:       # hidden * gelu(gate) in the projection's epilogue: the [B*HW, 8C] tensor never exists
            P.add_gemm([n], [(0, 1, 0, C // 64)], W_[f'{t}.ff1.wi'], 8 * C, f, bias=W_[f'{t}.ff1.bi'], reverse=self._rev(n, f),
                       label=f'{t}.ff.proj+geglu', geglu=True)
        else:
            g = self._act('ffg', B, H, H, 8 * C)
            P.add_gemm([n], [(0, 1, 0, C // 64)], W_[f'{t}.ff1.w'], 8 * C, g, bias=W_[f'{t}.ff1.b'], reverse=self._rev(n, g),
                       label=f'{t}.ff.proj')
            P.add_geglu(g, f, label=f'{t}.ff.geglu')
        t3 = self._act('t1', B, H, H, C)
        P.add_gemm([f], [(0, 1, 0, 4 * C // 64)], W_[f'{t}.ff2.w'], C, t3, bias=W_[f'{t}.ff2.b'], residual=t2,
                   reverse=self._rev(f, t3), label=f'{t}.ff.out')
        out = self._persist(f'{p}.out', H, H, C)
        P.add_gemm([t3], [(0, 1, 0, C // 64)], W_[f'{p}.proj_out.w'], C, out, bias=W_[f'{p}.proj_out.b'], residual=x,
                   gn_stats=self._new_stats(out), reverse=self._rev(t3, out), label=f'{p}.proj_out')
        self.block_out[p] = out
        return out

    # ------------------------------------------------------------------ whole network
    def _build_sd(self, eng: 'SDUNetEngine'):
        P, W_, cfg = self.plan, eng.w, eng.cfg
        B, H = self.B, self.H
        dev = self.x_in.device
        f32 = dict(device=dev, dtype=torch.float32)
        boc, lpb = cfg['block_out_channels'], cfg['layers_per_block']
        E = boc[0] * 4
        # timestep embedding MLP; only silu(emb) is consumed downstream (resnet.py: time_emb_proj(nonlinearity(temb)))
        e1 = torch.empty(1, E, **f32)
        self.temb_act = torch.empty(1, E, **f32)
        P.add_linear(self.emb_in, W_['time_embedding.linear_1.weight'], e1, bias=W_['time_embedding.linear_1.bias'], act=1,
                     label='time_embedding.linear_1+silu')
        P.add_linear(e1, W_['time_embedding.linear_2.weight'], self.temb_act, bias=W_['time_embedding.linear_2.bias'], act=1,
                     label='time_embedding.linear_2+silu')
        # all time_emb_proj layers at once, with each resnet's conv1 bias folded in -> per-resnet effective conv1 bias
        self.temb_bias = torch.empty(1, eng.temb_total, **f32)
        P.add_linear(self.temb_act, W_['temb_all.weight'], self.temb_bias, bias=W_['temb_all.bias'], label='time_emb_proj_all')

        c0 = boc[0]
        col = self._act('col', B, H, H, 64)
        P.add_im2col(self.x_in, col, label='conv_in.im2col')
        x = self._persist('conv_in', H, H, c0)
        P.add_gemm([col], [(0, 1, 0, 1)], W_['conv_in.w'], c0, x, bias=W_['conv_in.b'], alg_k=9 * cfg['in_channels'],
                   gn_stats=self._new_stats(x), reverse=self._rev(col, x), label='conv_in')
        self.block_out['conv_in'] = x
        skips = [x]
        res = H
        for i in range(len(boc)):
            for j in range(lpb):
                x = self._resnet(eng, f'down_blocks.{i}.resnets.{j}', [x])
                if cfg['cross_attn_down'][i]:
                    x = self._transformer(eng, f'down_blocks.{i}.attentions.{j}', x)
                skips.append(x)
            if i != len(boc) - 1:
                p = f'down_blocks.{i}.downsamplers.0.conv'
                res //= 2
                y = self._persist(p, res, res, boc[i])
                P.add_gemm([x], [(0, 9, 0, boc[i] // 64)], W_[f'{p}.w'], boc[i], y, bias=W_[f'{p}.b'], a_stride=[2],
                           gn_stats=self._new_stats(y), reverse=self._rev(x, y), label=p)
                self.block_out[p] = y
                x = y
                skips.append(x)
        x = self._resnet(eng, 'mid_block.resnets.0', [x])
        x = self._transformer(eng, 'mid_block.attentions.0', x)
        x = self._resnet(eng, 'mid_block.resnets.1', [x])
        cross_up = list(cfg['cross_attn_down'])[::-1]
        rev = boc[::-1]
        for i in range(len(boc)):
            for j in range(lpb + 1):
                x = self._resnet(eng, f'up_blocks.{i}.resnets.{j}', [x, skips.pop()])
                if cross_up[i]:
                    x = self._transformer(eng, f'up_blocks.{i}.attentions.{j}', x)
            if i != len(boc) - 1:
                p = f'up_blocks.{i}.upsamplers.0.conv'
                res *= 2
                y = self._persist(p, res, res, rev[i])
                if eng.fused_upsample:      # conv3x3(nearest_up2(x)) as four 2x2-tap phase launches over the low-res x
                    P.add_gemm([x], [(0, 9, 0, rev[i] // 64)], W_[f'{p}.wup'], rev[i], y, bias=W_[f'{p}.b'],
                               gn_stats=self._new_stats(y), reverse=self._rev(x, y), label=p, upsample2x=True)
                else:
                    u = self._act('up', B, res, res, rev[i])
                    P.add_upsample2x(x, u, label=f'up_blocks.{i}.upsample')
                    P.add_gemm([u], [(0, 9, 0, rev[i] // 64)], W_[f'{p}.w'], rev[i], y, bias=W_[f'{p}.b'],
                               gn_stats=self._new_stats(y), reverse=self._rev(u, y), label=p)
                self.block_out[p] = y
                x = y
        a = self._act('a0', B, H, H, c0)
        self._gn([x], c0, H, H, W_['conv_norm_out.weight'], W_['conv_norm_out.bias'], a, silu=True, label='conv_norm_out')
        P.add_gemm([a], [(0, 9, 0, c0 // 64)], W_['conv_out.w'], cfg['out_channels'], self.out, bias=W_['conv_out.b'],
                   reverse=self._rev(a), label='conv_out')


class SDUNetEngine:
    """Packed weights + cached plans.  `set_context(ctx_pair)` once per prompt, then `forward(x, t)`."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', use_graphs: bool = True, fused_geglu: bool = True,
                 fused_upsample: bool = True):
        from . import _lib
        _lib.lib()
        self.device = torch.device(device)
        self.fused_geglu = fused_geglu
        self.fused_upsample = fused_upsample
        if self.device.type != 'cuda':
            raise RuntimeError('SDUNetEngine requires a CUDA device (B200); there is no CPU fallback')
        self.use_graphs = use_graphs
        self.cfg = sd_config_from_state_dict(state_dict)
        for c in self.cfg['block_out_channels']:
            if c % 64:
                raise NotImplementedError('channel counts must be multiples of 64')
        self.w: Dict[str, torch.Tensor] = {}
        self.cout: Dict[str, int] = {}
        self.temb_off: Dict[str, int] = {}
        self.ctx_kv: Dict[str, torch.Tensor] = {}
        self.ctx_len = 0
        self._kvw: Dict[str, torch.Tensor] = {}
        self._pack(state_dict)
        self._plans: Dict[tuple, SDPlan] = {}

    # ------------------------------------------------------------------ weights
    def _pack(self, sd):
        dev, w, cfg = self.device, self.w, self.cfg
        heads = cfg['heads']
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        cpu = lambda t: t.detach().float().cpu()
        bf = lambda t: t.contiguous().to(ACT_DTYPE).to(dev)
        for k in ('time_embedding.linear_1.weight', 'time_embedding.linear_1.bias', 'time_embedding.linear_2.weight',
                  'time_embedding.linear_2.bias', 'conv_norm_out.weight', 'conv_norm_out.bias'):
            w[k] = f(sd[k])
        cin = cfg['in_channels']
        wp = torch.zeros(sd['conv_in.weight'].shape[0], 64, dtype=ACT_DTYPE)
        wp[:, :9 * cin] = _pack_conv(cpu(sd['conv_in.weight']))
        w['conv_in.w'], w['conv_in.b'] = wp.to(dev), f(sd['conv_in.bias'])
        wo = cpu(sd['conv_out.weight'])
        wp = torch.zeros(16, 9 * wo.shape[1], dtype=ACT_DTYPE)
        wp[:wo.shape[0]] = _pack_conv(wo)
        w['conv_out.w'], w['conv_out.b'] = wp.to(dev), f(sd['conv_out.bias'])

        resnets = sorted({k[:-len('.conv1.weight')] for k in sd if k.endswith('.conv1.weight')})
        temb_w, temb_b, off = [], [], 0
        for p in resnets:
            w1 = cpu(sd[f'{p}.conv1.weight'])
            cout = w1.shape[0]
            self.cout[p] = cout
            for nm in ('norm1', 'norm2'):
                w[f'{p}.{nm}.weight'], w[f'{p}.{nm}.bias'] = f(sd[f'{p}.{nm}.weight']), f(sd[f'{p}.{nm}.bias'])
            w[f'{p}.conv1.w'] = bf(_pack_conv(w1))            # conv1 reads the normalised concat: plain (tap, channel) order
            self.temb_off[p] = off
            temb_w.append(cpu(sd[f'{p}.time_emb_proj.weight']))
            temb_b.append(cpu(sd[f'{p}.time_emb_proj.bias']) + cpu(sd[f'{p}.conv1.bias']))
            off += cout
            w2 = _pack_conv(cpu(sd[f'{p}.conv2.weight']))
            if f'{p}.conv_shortcut.weight' in sd:            # [conv2 taps x cout | shortcut over the raw (concat) input]
                ws = cpu(sd[f'{p}.conv_shortcut.weight'])[:, :, 0, 0].to(ACT_DTYPE)
                w[f'{p}.conv2sc.w'] = torch.cat([w2, ws], dim=1).contiguous().to(dev)
                w[f'{p}.conv2sc.b'] = f(sd[f'{p}.conv2.bias']) + f(sd[f'{p}.conv_shortcut.bias'])
            else:
                w[f'{p}.conv2.w'], w[f'{p}.conv2.b'] = w2.to(dev), f(sd[f'{p}.conv2.bias'])
        self.temb_total = off
        w['temb_all.weight'] = torch.cat(temb_w, dim=0).contiguous().to(dev)
        w['temb_all.bias'] = torch.cat(temb_b, dim=0).contiguous().to(dev)

        for k in sd:
            if k.endswith(('downsamplers.0.conv.weight', 'upsamplers.0.conv.weight')):
                p = k[:-len('.weight')]
                w[f'{p}.b'] = f(sd[f'{p}.bias'])
                if 'upsamplers' in k and self.fused_upsample:
                    w[f'{p}.wup'] = pack_conv_up2(cpu(sd[k])).to(dev)
                else:
                    w[f'{p}.w'] = bf(_pack_conv(cpu(sd[k])))

        def pad_rows(m, hd, dp):          # [heads*hd, K] -> [heads*dp, K], zero rows between heads
            out = torch.zeros(heads * dp, m.shape[1])
            out.view(heads, dp, -1)[:, :hd] = m.view(heads, hd, -1)
            return out

        def pad_cols(m, hd, dp):          # [N, heads*hd] -> [N, heads*dp]
            out = torch.zeros(m.shape[0], heads * dp)
            out.view(m.shape[0], heads, dp)[:, :, :hd] = m.view(m.shape[0], heads, hd)
            return out

        for p in sorted({k[:-len('.proj_in.weight')] for k in sd if k.endswith('.proj_in.weight')}):
            C = sd[f'{p}.proj_in.weight'].shape[0]
            hd, dp = C // heads, _pad_dim(C // heads)
            w[f'{p}.norm.weight'], w[f'{p}.norm.bias'] = f(sd[f'{p}.norm.weight']), f(sd[f'{p}.norm.bias'])
            for nm in ('proj_in', 'proj_out'):
                w[f'{p}.{nm}.w'] = bf(cpu(sd[f'{p}.{nm}.weight'])[:, :, 0, 0])
                w[f'{p}.{nm}.b'] = f(sd[f'{p}.{nm}.bias'])
            t = f'{p}.transformer_blocks.0'
            for nm in ('norm1', 'norm2', 'norm3'):
                w[f'{t}.{nm}.weight'], w[f'{t}.{nm}.bias'] = f(sd[f'{t}.{nm}.weight']), f(sd[f'{t}.{nm}.bias'])
            a1, a2 = f'{t}.attn1', f'{t}.attn2'
            w[f'{a1}.qkv.w'] = bf(torch.cat([pad_rows(cpu(sd[f'{a1}.to_{x}.weight']), hd, dp) for x in 'qkv'], dim=0))
            w[f'{a2}.q.w'] = bf(pad_rows(cpu(sd[f'{a2}.to_q.weight']), hd, dp))
            self._kvw[a2] = bf(torch.cat([pad_rows(cpu(sd[f'{a2}.to_{x}.weight']), hd, dp) for x in 'kv'], dim=0))
            for a in (a1, a2):
                w[f'{a}.out.w'] = bf(pad_cols(cpu(sd[f'{a}.to_out.0.weight']), hd, dp))
                w[f'{a}.out.b'] = f(sd[f'{a}.to_out.0.bias'])
            if self.fused_geglu:      # rows regrouped as [64 hidden | 64 gate] for the fused GEGLU epilogue
                w[f'{t}.ff1.wi'] = bf(interleave_geglu(cpu(sd[f'{t}.ff.net.0.proj.weight'])))
                w[f'{t}.ff1.bi'] = f(interleave_geglu(sd[f'{t}.ff.net.0.proj.bias']))
            else:
                w[f'{t}.ff1.w'], w[f'{t}.ff1.b'] = bf(cpu(sd[f'{t}.ff.net.0.proj.weight'])), f(sd[f'{t}.ff.net.0.proj.bias'])
            w[f'{t}.ff2.w'], w[f'{t}.ff2.b'] = bf(cpu(sd[f'{t}.ff.net.2.weight'])), f(sd[f'{t}.ff.net.2.bias'])

    # ------------------------------------------------------------------ context
    def set_context(self, ctx_pair: torch.Tensor):
        """ctx_pair [2, T<=128, cross_dim] = encoder_hidden_states of [uncond, cond].  Projects K and V of every
        cross-attention layer once (attention_processor.py: to_k / to_v of encoder_hidden_states)."""
        if ctx_pair.dim() != 3 or ctx_pair.shape[0] != 2 or ctx_pair.shape[1] > CTX_ROWS:
            raise ValueError('ctx_pair must be [2, T <= 128, cross_attention_dim]')
        T, Dc = ctx_pair.shape[1], ctx_pair.shape[2]
        if Dc != self.cfg['cross_attention_dim'] or Dc % 64:
            raise ValueError('context width does not match the UNet (and must be a multiple of 64)')
        self.ctx_len = T
        ctx = torch.zeros(2, 1, CTX_ROWS, Dc, device=self.device, dtype=ACT_DTYPE)     # zero padded tokens
        ctx[:, 0, :T] = ctx_pair.to(device=self.device, dtype=ACT_DTYPE)
        plan = Plan()
        for name, wkv in self._kvw.items():
            kv = self.ctx_kv.get(name)
            if kv is None:
                kv = torch.empty(2, 1, CTX_ROWS, wkv.shape[0], device=self.device, dtype=ACT_DTYPE)
                self.ctx_kv[name] = kv
            plan.add_gemm([ctx], [(0, 1, 0, Dc // 64)], wkv, wkv.shape[0], kv, label=f'{name}.kv')
        plan.run()
        torch.cuda.synchronize(self.device)
        self._ctx = ctx

    # ------------------------------------------------------------------ forward
    def plan(self, B: int, H: int) -> SDPlan:
        if not self.ctx_kv:
            raise RuntimeError('call set_context(ctx_pair) before building a plan')
        if B % 2:
            raise ValueError('the UNet batch is [uncond half; cond half]: it must be even')
        key = (B, H)
        if key not in self._plans:
            self._plans[key] = SDPlan(self, B, H)
        return self._plans[key]

    def timestep_embedding(self, t) -> torch.Tensor:
        """embeddings.py get_timestep_embedding (flip_sin_to_cos=True, downscale_freq_shift=0) -> [1, C0] fp32."""
        dim = self.cfg['block_out_channels'][0]
        half = dim // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=self.device) / half
        emb = torch.as_tensor(t, device=self.device).reshape(1, 1).float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)

    def run(self, fp: SDPlan, t) -> torch.Tensor:
        """Run the plan on fp.x_in (already filled).  Returns fp.out (eps, fp32 NHWC)."""
        fp.emb_in.copy_(self.timestep_embedding(t))
        fp.plan.run()
        return fp.out

    def forward(self, x: torch.Tensor, t) -> torch.Tensor:
        """x fp32 NCHW [B,4,H,W] (rows [uncond half; cond half]) -> eps fp32 NCHW (view of the NHWC result)."""
        fp = self.plan(x.shape[0], x.shape[2])
        fp.x_in.copy_(x)
        return self.run(fp, t).permute(0, 3, 1, 2)

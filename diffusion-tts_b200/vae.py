"""Decode half of the SD `AutoencoderKL` as a static plan of sm_100a kernels (VAE decode inside the search
loop -- the reference decodes every candidate's Tweedie x0 to a 512x512 image before scoring it,
pipeline_stable_diffusion.py:1111-1118).

Replaces `AutoencoderKL.decode` (sd/diffusers/src/diffusers/models/autoencoders/autoencoder_kl.py:287-320) = `post_quant_conv`
+ `Decoder.forward` (vae.py:291-340): conv_in -> UNetMidBlock2D (ResnetBlock2D, one-head attention, ResnetBlock2D) ->
UpDecoderBlock2D x 4 (3 ResnetBlock2D each, nearest x2 + conv between levels) -> GroupNorm(32, eps 1e-6) -> SiLU -> conv_out.
Described by the state dict alone; activations are bf16 NHWC, images come out fp32 NHWC [B, 8h, 8w, 3].

Everything except the attention reuses the U-Net kernels: tcgen05 implicit-GEMM 3x3 convs with the GroupNorm statistics
fused into their epilogues (rows of 256 / 512 pixels are tiled as 128-pixel row segments), gn_finalize / gn_apply(+SiLU),
upsample2x, the 16-wide fp32 GEMM for conv_out.  `conv2 + conv_shortcut` share one accumulator where channels change.

The mid-block attention has ONE head of dimension C = 512 over L = h*w = 4096 tokens; a 128 x 512 fp32 accumulator alone
fills TMEM, so it runs unfused, per sample, on the GEMM kernel (0.07 of the decoder's 2.48 TFLOP):
    q, k     = 1x1 GEMMs over the whole batch (with bias)
    V^T      = W_v X^T      : GEMM with the WEIGHT matrix as the A operand and the sample's tokens as the B operand -> [C, L]
    S        = Q K^T        : GEMM, fp32 output [L, L]
    P        = softmax(S / sqrt(C))                                    (csrc/vae.cuh softmax_rows_kernel, fp32 -> bf16)
    O        = P V + b_v    : GEMM over the L keys; rows of P sum to 1, so the value bias is added after the product
    out      = to_out(O) + x
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from ._lib import ACT_DTYPE

from . import ops
from .ops import Plan
from .unet import ForwardPlan, _pack_conv


def vae_config_from_state_dict(sd: Dict[str, torch.Tensor]) -> dict:
    n_up = 1 + max(int(k.split('.')[2]) for k in sd if k.startswith('decoder.up_blocks.'))
    rev = [sd[f'decoder.up_blocks.{i}.resnets.0.conv1.weight'].shape[0] for i in range(n_up)]
    n_res = 1 + max(int(k.split('.')[4]) for k in sd if k.startswith('decoder.up_blocks.0.resnets.'))
    return dict(up_channels=rev, resnets_per_block=n_res, latent_channels=sd['decoder.conv_in.weight'].shape[1],
                top=sd['decoder.conv_in.weight'].shape[0], out_channels=sd['decoder.conv_out.weight'].shape[0])


class VAEPlan(ForwardPlan):
    """Buffers + kernel plan of the decoder for a fixed batch B of h x h latents."""

    def __init__(self, eng: 'VAEDecoderEngine', B: int, h: int):
        dev, cfg = eng.device, eng.cfg
        self.B, self.B_full, self.b_emb = B, B, 1
        self.h = h
        f32 = dict(device=dev, dtype=torch.float32)
        self.x_in = torch.zeros(B, cfg['latent_channels'], h, h, **f32)                 # z / scaling_factor, fp32 NCHW
        self.z = torch.zeros_like(self.x_in)                                            # after post_quant_conv
        H_out = h * 2 ** (len(cfg['up_channels']) - 1)
        self.out = torch.empty(B, H_out, H_out, cfg['out_channels'], **f32)             # image, fp32 NHWC
        self.plan = Plan()
        self._scratch: Dict[str, torch.Tensor] = {}
        self._full: Dict[str, torch.Tensor] = {}
        self.block_out: Dict[str, torch.Tensor] = {}
        self._stats: Dict[tuple, torch.Tensor] = {}
        self._dir: Dict[tuple, bool] = {}
        self.fused_gn_stats = True
        self.alternate_walk = True
        self.n_lanes, self._lane = 1, None
        self._eps = 1e-6                                                                # resnet_eps / norm eps (vae.py:247-262)
        self._flip = 0
        self.keep_taps = eng.keep_taps
        self._build(eng)
        if os.environ.get('B200NS_PDL') is None and self.B_full <= 4:
            self.plan.set_pdl(1)          # small batches are launch-latency bound: overlap each kernel's prologue with its predecessor
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            self.plan.instantiate_graph()

    @staticmethod
    def _num_groups(C: int) -> int:
        return 32                                                                       # norm_num_groups

    def _out_buf(self, name: str, H: int, C: int) -> torch.Tensor:
        """Block outputs ping-pong between two scratch buffers (a block never writes the buffer it reads); with
        keep_taps (parity tests) every block output gets its own tensor instead."""
        if self.keep_taps:
            t = self._persist(name, H, H, C)
            self.block_out[name] = t
            return t
        self._flip ^= 1
        return self._act(f'x{self._flip}', self.B, H, H, C)

    def _stats_for(self, t: torch.Tensor):
        return self._new_stats(t) if self.keep_taps else self._new_stats(t, key=f'xs{self._flip}')

    def _resnet(self, eng, p: str, x: torch.Tensor) -> torch.Tensor:
        P, W_ = self.plan, eng.w
        B, H, _, cin = x.shape
        cout = eng.cout[p]
        a0 = self._act('a0', B, H, H, cin)
        self._gn([x], cin, H, H, W_[f'{p}.norm1.weight'], W_[f'{p}.norm1.bias'], a0, silu=True, label=f'{p}.norm1')
        hbuf = self._act('h', B, H, H, cout)
        P.add_gemm([a0], [(0, 9, 0, cin // 64)], W_[f'{p}.conv1.w'], cout, hbuf, bias=W_[f'{p}.conv1.b'],
                   gn_stats=self._new_stats(hbuf, 'h_stats'), reverse=self._rev(a0, hbuf), label=f'{p}.conv1')
        a1 = self._act('a1', B, H, H, cout)
        self._gn([hbuf], cout, H, H, W_[f'{p}.norm2.weight'], W_[f'{p}.norm2.bias'], a1, silu=True, label=f'{p}.norm2')
        out = self._out_buf(p, H, cout)
        if f'{p}.conv2sc.w' in W_:       # conv2 over the normalised hidden + 1x1 shortcut over the raw input, one accumulator
            P.add_gemm([a1, x], [(0, 9, 0, cout // 64), (1, 1, 0, cin // 64)], W_[f'{p}.conv2sc.w'], cout, out,
                       bias=W_[f'{p}.conv2sc.b'], gn_stats=self._stats_for(out), reverse=self._rev(a1, out),
                       label=f'{p}.conv2+shortcut')
        else:
            P.add_gemm([a1], [(0, 9, 0, cout // 64)], W_[f'{p}.conv2.w'], cout, out, bias=W_[f'{p}.conv2.b'], residual=x,
                       gn_stats=self._stats_for(out), reverse=self._rev(a1, out), label=f'{p}.conv2')
        return out

    def _attention(self, eng, p: str, x: torch.Tensor) -> torch.Tensor:
        P, W_ = self.plan, eng.w
        B, H, _, C = x.shape
        L = H * H
        dev = x.device
        xn = self._act('a0', B, H, H, C)
        self._gn([x], C, H, H, W_[f'{p}.group_norm.weight'], W_[f'{p}.group_norm.bias'], xn, silu=False, label=f'{p}.group_norm')
        q, k = self._act('h', B, H, H, C), self._act('a1', B, H, H, C)
        P.add_gemm([xn], [(0, 1, 0, C // 64)], W_[f'{p}.q.w'], C, q, bias=W_[f'{p}.q.b'], label=f'{p}.to_q')
        P.add_gemm([xn], [(0, 1, 0, C // 64)], W_[f'{p}.k.w'], C, k, bias=W_[f'{p}.k.b'], label=f'{p}.to_k')
        o = self._act('att_o', B, H, H, C)
        S = self._buf('att_s', L * L, torch.float32)[:L * L].view(1, H, H, L)
        Pm = self._act('att_p', 1, H, H, L)
        wrows = 128 if C % 128 == 0 else 64
        vt = self._act('att_vt', 1, C // wrows, wrows, L)                      # V^T [C, L] as an "image" of C pixels
        wv_img = W_[f'{p}.v.w'].view(1, C // wrows, wrows, C)                   # rows = output features, channels = inputs
        for s in range(B):
            tok = xn[s].reshape(L, C)
            P.add_gemm([wv_img], [(0, 1, 0, C // 64)], tok, L, vt, label=f'{p}.vT[{s}]')
            P.add_gemm([q[s:s + 1]], [(0, 1, 0, C // 64)], k[s].reshape(L, C), L, S, label=f'{p}.qk[{s}]')
            P.add_softmax_rows(S.view(L, L), Pm.view(L, L), 1.0 / math.sqrt(C), label=f'{p}.softmax[{s}]')
            P.add_gemm([Pm], [(0, 1, 0, L // 64)], vt.view(C, L), C, o[s:s + 1], bias=W_[f'{p}.v.b'], label=f'{p}.pv[{s}]')
        out = self._out_buf(p, H, C)
        P.add_gemm([o], [(0, 1, 0, C // 64)], W_[f'{p}.out.w'], C, out, bias=W_[f'{p}.out.b'], residual=x,
                   gn_stats=self._stats_for(out), label=f'{p}.to_out')
        return out

    def _build(self, eng: 'VAEDecoderEngine'):
        cfg, P, W_ = eng.cfg, self.plan, eng.w
        B, H = self.B, self.h
        top = cfg['top']
        col = self._act('col', B, H, H, 64)
        P.add_im2col(self.z, col, label='decoder.conv_in.im2col')
        x = self._out_buf('decoder.conv_in', H, top)
        P.add_gemm([col], [(0, 1, 0, 1)], W_['decoder.conv_in.w'], top, x, bias=W_['decoder.conv_in.b'],
                   alg_k=9 * cfg['latent_channels'], gn_stats=self._stats_for(x), reverse=self._rev(col, x),
                   label='decoder.conv_in')
        x = self._resnet(eng, 'decoder.mid_block.resnets.0', x)
        x = self._attention(eng, 'decoder.mid_block.attentions.0', x)
        x = self._resnet(eng, 'decoder.mid_block.resnets.1', x)
        res = H
        ups = cfg['up_channels']
        for i, c in enumerate(ups):
            for j in range(cfg['resnets_per_block']):
                x = self._resnet(eng, f'decoder.up_blocks.{i}.resnets.{j}', x)
            if i != len(ups) - 1:
                p = f'decoder.up_blocks.{i}.upsamplers.0.conv'
                res *= 2
                y = self._out_buf(p, res, c)
                if eng.fused_upsample:      # conv3x3(nearest_up2(x)) as four 2x2-tap phase launches over the low-res x
                    P.add_gemm([x], [(0, 9, 0, c // 64)], W_[f'{p}.wup'], c, y, bias=W_[f'{p}.b'], gn_stats=self._stats_for(y),
                               reverse=self._rev(x, y), label=p, upsample2x=True)
                else:
                    u = self._act('up', B, res, res, c)
                    P.add_upsample2x(x, u, label=f'decoder.up_blocks.{i}.upsample')
                    P.add_gemm([u], [(0, 9, 0, c // 64)], W_[f'{p}.w'], c, y, bias=W_[f'{p}.b'], gn_stats=self._stats_for(y),
                               reverse=self._rev(u, y), label=p)
                x = y
        c0 = ups[-1]
        a = self._act('a0', B, res, res, c0)
        self._gn([x], c0, res, res, W_['decoder.conv_norm_out.weight'], W_['decoder.conv_norm_out.bias'], a, silu=True,
                 label='decoder.conv_norm_out')
        P.add_gemm([a], [(0, 9, 0, c0 // 64)], W_['decoder.conv_out.w'], cfg['out_channels'], self.out,
                   bias=W_['decoder.conv_out.b'], reverse=self._rev(a), label='decoder.conv_out')


class VAEDecoderEngine:
    """Packed weights + cached plans.  `decode(z)` = AutoencoderKL.decode(z) for z already divided by scaling_factor;
    `decode_latents(x0)` applies the 1 / scaling_factor of the pipeline (pipeline_stable_diffusion.py:1112)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', use_graphs: bool = True, scaling_factor: float = 0.18215,
                 keep_taps: bool = False, fused_upsample: bool = True):
        from . import _lib
        _lib.lib()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('VAEDecoderEngine requires a CUDA device (B200); there is no CPU fallback')
        self.use_graphs, self.keep_taps, self.scaling_factor = use_graphs, keep_taps, scaling_factor
        self.fused_upsample = fused_upsample
        self.cfg = vae_config_from_state_dict(state_dict)
        for c in self.cfg['up_channels'] + [self.cfg['top']]:
            if c % 64:
                raise NotImplementedError('channel counts must be multiples of 64')
        self.w: Dict[str, torch.Tensor] = {}
        self.cout: Dict[str, int] = {}
        self._pack(state_dict)
        self._plans: Dict[tuple, VAEPlan] = {}

    def _pack(self, sd):
        dev, w = self.device, self.w
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        cpu = lambda t: t.detach().float().cpu()
        bf = lambda t: t.contiguous().to(ACT_DTYPE).to(dev)
        w['post_quant.w'] = f(sd['post_quant_conv.weight'][:, :, 0, 0])
        w['post_quant.b'] = f(sd['post_quant_conv.bias'])
        for k in ('decoder.conv_norm_out.weight', 'decoder.conv_norm_out.bias'):
            w[k] = f(sd[k])
        cin = self.cfg['latent_channels']
        wp = torch.zeros(self.cfg['top'], 64, dtype=ACT_DTYPE)
        wp[:, :9 * cin] = _pack_conv(cpu(sd['decoder.conv_in.weight']))
        w['decoder.conv_in.w'], w['decoder.conv_in.b'] = wp.to(dev), f(sd['decoder.conv_in.bias'])
        wo = cpu(sd['decoder.conv_out.weight'])
        wp = torch.zeros(16, 9 * wo.shape[1], dtype=ACT_DTYPE)
        wp[:wo.shape[0]] = _pack_conv(wo)
        w['decoder.conv_out.w'], w['decoder.conv_out.b'] = wp.to(dev), f(sd['decoder.conv_out.bias'])
        for p in sorted({k[:-len('.conv1.weight')] for k in sd if k.startswith('decoder.') and k.endswith('.conv1.weight')}):
            w1 = cpu(sd[f'{p}.conv1.weight'])
            self.cout[p] = w1.shape[0]
            for nm in ('norm1', 'norm2'):
                w[f'{p}.{nm}.weight'], w[f'{p}.{nm}.bias'] = f(sd[f'{p}.{nm}.weight']), f(sd[f'{p}.{nm}.bias'])
            w[f'{p}.conv1.w'], w[f'{p}.conv1.b'] = bf(_pack_conv(w1)), f(sd[f'{p}.conv1.bias'])
            w2 = _pack_conv(cpu(sd[f'{p}.conv2.weight']))
            if f'{p}.conv_shortcut.weight' in sd:
                ws = cpu(sd[f'{p}.conv_shortcut.weight'])[:, :, 0, 0].to(ACT_DTYPE)
                w[f'{p}.conv2sc.w'] = torch.cat([w2, ws], dim=1).contiguous().to(dev)
                w[f'{p}.conv2sc.b'] = f(sd[f'{p}.conv2.bias']) + f(sd[f'{p}.conv_shortcut.bias'])
            else:
                w[f'{p}.conv2.w'], w[f'{p}.conv2.b'] = w2.to(dev), f(sd[f'{p}.conv2.bias'])
        for p in sorted({k[:-len('.conv.weight')] for k in sd if k.startswith('decoder.') and k.endswith('upsamplers.0.conv.weight')}):
            w[f'{p}.conv.b'] = f(sd[f'{p}.conv.bias'])
            if self.fused_upsample:
                w[f'{p}.conv.wup'] = ops.pack_conv_up2(cpu(sd[f'{p}.conv.weight'])).to(dev)
            else:
                w[f'{p}.conv.w'] = bf(_pack_conv(cpu(sd[f'{p}.conv.weight'])))
        a = 'decoder.mid_block.attentions.0'
        w[f'{a}.group_norm.weight'], w[f'{a}.group_norm.bias'] = f(sd[f'{a}.group_norm.weight']), f(sd[f'{a}.group_norm.bias'])
        for nm, src in (('q', 'to_q'), ('k', 'to_k'), ('v', 'to_v'), ('out', 'to_out.0')):
            w[f'{a}.{nm}.w'], w[f'{a}.{nm}.b'] = bf(cpu(sd[f'{a}.{src}.weight'])), f(sd[f'{a}.{src}.bias'])

    def plan(self, B: int, h: int) -> VAEPlan:
        key = (B, h)
        if key not in self._plans:
            self._plans[key] = VAEPlan(self, B, h)
        return self._plans[key]

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z fp32 NCHW [B, 4, h, w] -> image fp32 NHWC [B, 8h, 8w, 3] (a view of the plan's static output buffer)."""
        if not z.is_cuda:
            raise RuntimeError('VAEDecoderEngine.decode needs a CUDA tensor (no CPU fallback)')
        fp = self.plan(z.shape[0], z.shape[2])
        fp.x_in.copy_(z)
        ops.post_quant(fp.x_in, self.w['post_quant.w'], self.w['post_quant.b'], out=fp.z)
        fp.plan.run()
        return fp.out

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        return self.decode(latents.to(torch.float32) / self.scaling_factor)


class DecodedImageScorer:
    """`scores = scorer(x0)` for the SD search drivers (`decode` = identity): decodes each candidate's Tweedie x0 to an image
    (pipeline_stable_diffusion.py:1111-1118: `vae.decode(pred_x0 / scaling_factor)`, `(image*127.5+128).clip(0,255).to(uint8)`)
    and scores it -- in chunks of `chunk` candidates, so a beam step with hundreds of candidates needs the activations of
    only `chunk` 512x512 decodes at a time.  `score_function` None or a scorer with `latent_fused` (BrightnessScorer):
    the RGB luminance comes from exact integer channel sums of the quantised image and no uint8 image is materialised;
    any other scorer receives `score_function(images_uint8 [M,3,H,W], prompts, timesteps)` like the reference."""

    def __init__(self, vae: VAEDecoderEngine, score_function=None, prompt: str = '', chunk: int = 16):
        self.vae, self.fn, self.prompt, self.chunk = vae, score_function, prompt, chunk
        self.fused = score_function is None or getattr(score_function, 'latent_fused', False)
        self.decoded = 0

    @torch.no_grad()
    def __call__(self, x0: torch.Tensor) -> torch.Tensor:
        R = x0.shape[0]
        out = torch.empty(R, device=x0.device, dtype=torch.float32)
        for lo in range(0, R, self.chunk):
            hi = min(R, lo + self.chunk)
            img = self.vae.decode_latents(x0[lo:hi])
            sums, u8 = ops.image_sums(img, want_u8=not self.fused)
            if self.fused:
                out[lo:hi] = ops.brightness_from_sums(sums, img.shape[3], img.shape[1] * img.shape[2])
            else:
                s = self.fn(u8, [self.prompt] * (hi - lo), torch.zeros(hi - lo, device=x0.device))
                out[lo:hi] = torch.as_tensor(s).to(device=x0.device, dtype=torch.float32).reshape(-1)
            self.decoded += hi - lo
        return out

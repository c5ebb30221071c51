"""ImageNet classifier scorer on the B200 engine.

The reference scores candidates with OpenAI's ADM 64x64 classifier -- `EncoderUNetModel`
(edm/unet.py:701-912) configured by `ImageNetScorer.create_classifier` (edm/scorers.py:101-140) --
and returns the softmax PROBABILITY of the target class (edm/scorers.py:162-172).  In the reference it
runs on the CPU (the scorer is never moved to the device, edm/scorers.py:144,156).  Here the torso
(scale-shift ResBlocks, AttentionBlocks) reuses the U-Net engine's kernels -- a ResBlock is exactly a
UNetBlock with skip_scale 1 (edm/unet.py:254-274 vs edm/training/networks.py:166-187) -- and the
CLIP-style attention pool is evaluated only for the one token whose output is used.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from ._lib import ACT_DTYPE

from .ops import Plan
from .scorers import Scorer
from .unet import Block, ForwardPlan, NetConfig, _pack_conv


def derive_classifier_layout(sd: Dict[str, torch.Tensor]):
    """Blocks of input_blocks[1:] + middle_block from state-dict keys: [(res_prefix, attn_prefix|None, cin, cout, down)].
    A ResBlock down-samples iff it is the last block of its level and keeps the channel count; the
    checkpoint does not store that flag, so it is recovered from the spatial size of the attention pool."""
    idxs = sorted({int(k.split('.')[1]) for k in sd if k.startswith('input_blocks.')})
    blocks = []
    for i in idxs[1:]:
        p = f'input_blocks.{i}'
        w = sd[f'{p}.0.in_layers.2.weight']
        blocks.append([f'{p}.0', f'{p}.1' if f'{p}.1.qkv.weight' in sd else None, w.shape[1], w.shape[0], False])
    return blocks


class ClassifierPlan(ForwardPlan):
    """Buffers + kernel plan of the classifier for a fixed batch."""

    def __init__(self, eng: 'ClassifierEngine', B: int):
        dev = eng.device
        self.B, self.b_emb = B, 1
        self.images = torch.zeros(B, 3, eng.image_size, eng.image_size, device=dev, dtype=torch.uint8)
        self.x_in = torch.zeros(B, 3, eng.image_size, eng.image_size, device=dev, dtype=torch.float32)
        self.target = torch.zeros(B, device=dev, dtype=torch.int64)
        self.t_emb = torch.zeros(1, eng.model_channels, device=dev, dtype=torch.float32)
        self.scores = torch.empty(B, device=dev, dtype=torch.float32)
        self.plan = Plan()
        self._scratch: Dict[str, torch.Tensor] = {}
        self.block_out: Dict[str, torch.Tensor] = {}
        self._stats: Dict[int, torch.Tensor] = {}
        self.fused_gn_stats = True
        self._dir: Dict[tuple, bool] = {}
        self.alternate_walk = True
        self._full: Dict[str, torch.Tensor] = {}
        self.B_full, self.n_lanes, self._lane = B, 1, None
        self._eps = 1e-5
        self._build_classifier(eng)
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            self.plan.instantiate_graph()

    @staticmethod
    def _num_groups(C: int) -> int:
        return 32                         # GroupNorm32(32, channels) (edm/nn_utils.py:93-100)

    def _build_classifier(self, eng: 'ClassifierEngine'):
        P, W_, B = self.plan, eng.w, self.B
        dev = self.x_in.device
        f32 = dict(device=dev, dtype=torch.float32)
        E = eng.emb_channels
        # time embedding -> SiLU -> all per-block (scale, shift) linears at once (edm/unet.py:262-270)
        t0 = torch.empty(1, E, **f32)
        self.emb = torch.empty(1, E, **f32)
        P.add_linear(self.t_emb, W_['time_embed.0.weight'], t0, bias=W_['time_embed.0.bias'], act=1, label='time_embed.0')
        P.add_linear(t0, W_['time_embed.2.weight'], self.emb, bias=W_['time_embed.2.bias'], act=1, label='time_embed.2+silu')
        self.film = torch.empty(1, eng.affine_total, **f32)
        P.add_linear(self.emb, W_['affine_all.weight'], self.film, bias=W_['affine_all.bias'], label='emb_layers_all')
        # input: uint8 -> [0,1] fp32 -> 3x3 conv via im2col
        P.add_u8_to_f32(self.images, self.x_in, label='u8_to_unit')
        H = eng.image_size
        col = self._act('col', B, H, H, 64)
        P.add_im2col(self.x_in, col, label='input_conv.im2col')
        c0 = eng.blocks[0].cin
        x = torch.empty(B, H, H, c0, device=dev, dtype=ACT_DTYPE)
        P.add_gemm([col], [(0, 1, 0, 1)], W_['input_conv.w'], c0, x, bias=W_['input_conv.b'], alg_k=27,
                   gn_stats=self._new_stats(x), label='input_conv')
        for blk in eng.blocks:
            x = self._block(eng, blk, [x])
            self.block_out[blk.name] = x
        # head: GroupNorm32 -> SiLU -> attention pool (token 0) -> softmax -> target probability
        Hs, C = x.shape[1], x.shape[3]
        T = Hs * Hs
        act = self._act('a0', B, Hs, Hs, C)
        self._gn([x], C, Hs, Hs, W_['out.0.weight'], W_['out.0.bias'], act, silu=True, label='out.norm')
        tok = self._act('a1', B, Hs, Hs, C)
        tok0 = torch.empty(B, C, **f32)
        P.add_pool_tokens(act.view(B, T, C), W_['pool.pos'], tok.view(B, T, C), tok0, B, T, C)
        kv = self._act('qkv', B, Hs, Hs, 2 * C)
        P.add_gemm([tok], [(0, 1, 0, C // 64)], W_['pool.kv.w'], 2 * C, kv, bias=W_['pool.kv.b'], label='pool.kv_proj')
        qkv0 = torch.empty(B, 3 * C, **f32)
        P.add_linear(tok0, W_['pool.qkv.weight'], qkv0, bias=W_['pool.qkv.bias'], label='pool.qkv_proj_token0')
        pooled = torch.empty(B, C, **f32)
        P.add_pool_attention(qkv0, kv.view(B, T, 2 * C), pooled, B, T, C)
        self.logits = torch.empty(B, eng.num_classes, **f32)
        P.add_linear(pooled, W_['pool.c_proj.weight'], self.logits, bias=W_['pool.c_proj.bias'], label='pool.c_proj')
        P.add_softmax_gather(self.logits, self.target, self.scores)


class ClassifierEngine:
    """Packed weights + cached plans of the classifier.  `scores(images_u8, target_idx)` -> prob[M]."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', use_graphs: bool = True):
        from . import _lib
        _lib.lib()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('ClassifierEngine requires a CUDA device (B200); there is no CPU fallback')
        self.use_graphs = use_graphs
        sd = state_dict
        self.model_channels = sd['time_embed.0.weight'].shape[1]
        self.emb_channels = sd['time_embed.0.weight'].shape[0]
        self.num_classes = sd['out.2.c_proj.weight'].shape[0]
        pos = sd['out.2.positional_embedding']
        self.pool_tokens = pos.shape[1] - 1
        layout = derive_classifier_layout(sd)
        # number of down-sampling ResBlocks = log2(image / pool side); they are the channel-preserving
        # blocks without attention that sit last in their level.  image_size is not stored in the
        # checkpoint: it is fixed by the caller's images (ImageNetScorer: 64).
        self.cfg = NetConfig(model_type='EncoderUNetModel', img_resolution=0, in_channels=3, out_channels=self.num_classes,
                             label_dim=0, noise_channels=self.model_channels, emb_channels=self.emb_channels,
                             adaptive_scale=True, skip_scale=1.0, eps=1e-5)
        self._layout = layout
        self._sd = sd
        self.w: Dict[str, torch.Tensor] = {}
        self.affine_off: Dict[str, int] = {}
        self.blocks: List[Block] = []
        self.image_size = None
        self._plans: Dict[int, ClassifierPlan] = {}

    def _finalize(self, image_size: int):
        """Resolve which ResBlocks down-sample (needs the image size), then pack the weights."""
        if self.image_size is not None:
            if image_size != self.image_size:
                raise ValueError(f'classifier was set up for {self.image_size}x{self.image_size} images')
            return
        self.image_size = image_size
        sd, dev, w = self._sd, self.device, self.w
        side = int(round(self.pool_tokens ** 0.5))
        n_down = int(round(np.log2(image_size / side)))
        # candidates for down blocks: cin == cout, no attention, followed by a block (edm/unet.py:775-796)
        layout = self._layout
        level_ends = [i for i, (rp, ap, cin, cout, _) in enumerate(layout)
                      if ap is None and cin == cout and i + 1 < len(layout) and
                      (layout[i + 1][2] == cout) and self._is_level_end(i)]
        if len(level_ends) < n_down:
            raise ValueError('cannot locate the down-sampling ResBlocks of the classifier')
        for i in level_ends[:n_down]:
            layout[i][4] = True
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        for k in ('time_embed.0.weight', 'time_embed.0.bias', 'time_embed.2.weight', 'time_embed.2.bias', 'out.0.weight',
                  'out.0.bias'):
            w[k] = f(sd[k])
        wi = sd['input_blocks.0.0.weight'].detach().float().cpu()
        wp = torch.zeros(wi.shape[0], 64, dtype=ACT_DTYPE)
        wp[:, :27] = _pack_conv(wi)
        w['input_conv.w'], w['input_conv.b'] = wp.to(dev), f(sd['input_blocks.0.0.bias'])
        res = image_size
        aff_w, aff_b, off = [], [], 0
        all_blocks = [tuple(b) for b in layout] + [('middle_block.0', 'middle_block.1', None, None, False),
                                                   ('middle_block.2', None, None, None, False)]
        for rp, ap, cin, cout, down in all_blocks:
            if cin is None:
                cw = sd[f'{rp}.in_layers.2.weight']
                cin, cout = cw.shape[1], cw.shape[0]
            if down:
                res //= 2
            n = rp
            blk = Block(n, 'block', cin, cout, res, up=False, down=down, attention=ap is not None,
                        heads=(cout // 64) if ap is not None else 0, skip_conv=f'{rp}.skip_connection.weight' in sd)
            self.blocks.append(blk)
            w[f'{n}.norm0.weight'], w[f'{n}.norm0.bias'] = f(sd[f'{rp}.in_layers.0.weight']), f(sd[f'{rp}.in_layers.0.bias'])
            w[f'{n}.norm1.weight'], w[f'{n}.norm1.bias'] = f(sd[f'{rp}.out_layers.0.weight']), f(sd[f'{rp}.out_layers.0.bias'])
            w[f'{n}.conv0.w'] = _pack_conv(sd[f'{rp}.in_layers.2.weight'].detach().float().cpu()).to(dev)
            w[f'{n}.conv0.b'] = f(sd[f'{rp}.in_layers.2.bias'])
            w1 = _pack_conv(sd[f'{rp}.out_layers.3.weight'].detach().float().cpu())
            if blk.skip_conv:
                ws = sd[f'{rp}.skip_connection.weight'].detach().float().cpu()[:, :, 0, 0].to(ACT_DTYPE)
                w[f'{n}.conv1skip.w'] = torch.cat([w1, ws], dim=1).contiguous().to(dev)
                w[f'{n}.conv1skip.b'] = f(sd[f'{rp}.out_layers.3.bias']) + f(sd[f'{rp}.skip_connection.bias'])
            else:
                w[f'{n}.conv1.w'], w[f'{n}.conv1.b'] = w1.to(dev), f(sd[f'{rp}.out_layers.3.bias'])
            if ap is not None:
                C, heads = cout, cout // 64
                w[f'{n}.norm2.weight'], w[f'{n}.norm2.bias'] = f(sd[f'{ap}.norm.weight']), f(sd[f'{ap}.norm.bias'])
                # QKVAttentionLegacy channel order (head, {q,k,v}, d) (edm/unet.py:365) -> [Q | K | V], head-major
                wq = sd[f'{ap}.qkv.weight'].detach().float().cpu()[:, :, 0].reshape(heads, 3, 64, C).permute(1, 0, 2, 3)
                bq = sd[f'{ap}.qkv.bias'].detach().float().cpu().reshape(heads, 3, 64).permute(1, 0, 2)
                w[f'{n}.qkv.w'] = wq.reshape(3 * C, C).contiguous().to(ACT_DTYPE).to(dev)
                w[f'{n}.qkv.b'] = bq.reshape(3 * C).contiguous().to(dev)
                w[f'{n}.proj.w'] = sd[f'{ap}.proj_out.weight'].detach().float().cpu()[:, :, 0].contiguous().to(
                    ACT_DTYPE).to(dev)
                w[f'{n}.proj.b'] = f(sd[f'{ap}.proj_out.bias'])
            self.affine_off[n] = off
            aff_w.append(sd[f'{rp}.emb_layers.1.weight'].detach().float().cpu())
            aff_b.append(sd[f'{rp}.emb_layers.1.bias'].detach().float().cpu())
            off += aff_w[-1].shape[0]
        self.affine_total = off
        w['affine_all.weight'] = torch.cat(aff_w, dim=0).contiguous().to(dev)
        w['affine_all.bias'] = torch.cat(aff_b, dim=0).contiguous().to(dev)
        C = self.blocks[-1].cout
        w['pool.pos'] = f(sd['out.2.positional_embedding'])
        wqkv = sd['out.2.qkv_proj.weight'].detach().float().cpu()[:, :, 0]          # rows [q | k | v] (edm/unet.py:398)
        w['pool.qkv.weight'], w['pool.qkv.bias'] = wqkv.contiguous().to(dev), f(sd['out.2.qkv_proj.bias'])
        w['pool.kv.w'] = wqkv[C:].contiguous().to(ACT_DTYPE).to(dev)
        w['pool.kv.b'] = f(sd['out.2.qkv_proj.bias'])[C:].contiguous()
        w['pool.c_proj.weight'] = sd['out.2.c_proj.weight'].detach().float()[:, :, 0].contiguous().to(dev)
        w['pool.c_proj.bias'] = f(sd['out.2.c_proj.bias'])
        self._sd = None

    def _is_level_end(self, i: int) -> bool:
        """Block i closes a level if the next block changes the channel count or starts with attention on a
        different width -- in the ADM classifier the only channel-preserving, attention-free block that is
        directly followed by a widening block is the down-sampling ResBlock (edm/unet.py:775-796)."""
        layout = self._layout
        nxt = layout[i + 1]
        return nxt[3] != nxt[2] or (layout[i][1] is None and nxt[1] is not None)

    def plan(self, B: int) -> ClassifierPlan:
        if B not in self._plans:
            self._plans[B] = ClassifierPlan(self, B)
        return self._plans[B]

    def timestep_embedding(self, t: float) -> torch.Tensor:
        """edm/nn_utils.py:103-121 for one timestep."""
        half = self.model_channels // 2
        freqs = torch.exp(-np.log(10000) * torch.arange(0, half, dtype=torch.float32, device=self.device) / half)
        args = torch.tensor([[float(t)]], device=self.device) * freqs[None]
        return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)

    @torch.no_grad()
    def scores(self, images_u8: torch.Tensor, target: torch.Tensor, timestep: float = 0.0) -> torch.Tensor:
        B, _, H, W = images_u8.shape
        self._finalize(H)
        cp = self.plan(B)
        cp.images.copy_(images_u8)
        cp.target.copy_(target)
        cp.t_emb.copy_(self.timestep_embedding(timestep))
        cp.plan.run()
        return cp.scores.clone()


class ImageNetScorer(Scorer):
    """Same call protocol as the reference's ImageNetScorer (edm/scorers.py:143-174): uint8 images [M,3,H,W],
    class labels (one-hot [M,K] or indices [M]), timesteps [M] -> probability of the target class.
    `state_dict` is the classifier checkpoint (the reference downloads `64x64_classifier.pt`, which is
    unreachable offline; any dict with EncoderUNetModel's keys works)."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, dtype=torch.float32, device='cuda'):
        super().__init__(dtype)
        if state_dict is None:
            raise RuntimeError('ImageNetScorer needs the classifier state dict (the pretrained 64x64_classifier.pt '
                               'cannot be downloaded offline): pass state_dict=torch.load(path)')
        self.engine = ClassifierEngine(state_dict, device=device)
        self.device = torch.device(device)

    @torch.no_grad()
    def __call__(self, images, class_labels, timesteps=None):
        if not isinstance(images, torch.Tensor) or images.dtype != torch.uint8 or images.dim() != 4:
            raise TypeError('B200 ImageNetScorer scores uint8 [M,3,H,W] images (edm/main.py:827)')
        images = images.to(self.device).contiguous()
        class_labels = class_labels.to(self.device)
        target = torch.argmax(class_labels, dim=1) if class_labels.dim() > 1 else class_labels
        t = 0.0
        if getattr(timesteps, '_b200_uniform_value', None) is not None:      # the search loop's own zeros: no host sync
            t = float(timesteps._b200_uniform_value)
        elif timesteps is not None and torch.is_tensor(timesteps) and timesteps.numel():
            if not bool((timesteps == timesteps.flatten()[0]).all()):
                raise NotImplementedError('per-sample timesteps are not used by the search path (always zeros)')
            t = float(timesteps.flatten()[0])
        return self.engine.scores(images, target.to(torch.int64).contiguous(), t).to(self.dtype)

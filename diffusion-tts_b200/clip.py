"""CLIP scorer on B200: the reference's `CLIPScorer` (sd/scorers.py:149-213) =
`CLIPProcessor(images)` -> `CLIPModel.get_image_features` -> cosine similarity with `get_text_features(prompt)`.

What runs per candidate image (the hot part) is native:
  * `CLIPImageProcessor`'s PIL path -- bicubic resize of the shortest edge to 224, centre crop, /255, normalise -- in
    integer arithmetic, bit-exact against Pillow, written straight into the patch-embedding GEMM's A matrix
    (csrc/clip.cuh; the resampling coefficients are computed here exactly as Resample.c:precompute_coeffs does);
  * the ViT tower (modeling_clip.py CLIPVisionTransformer) on the tcgen05 plan ops: patch embedding = one GEMM whose
    residual input carries class + position embeddings, pre-LN, per layer LN -> fused QKV GEMM -> flash attention over
    the 257 real tokens (rows padded to 384 per image, padded keys masked by `kv_len`) -> out-proj GEMM + residual ->
    LN -> fc1 GEMM with quick_gelu in the epilogue -> fc2 GEMM + residual; class-token pooling + post-LN + projection
    + cosine in fp32.  One CUDA graph per batch size.
The text embedding does not depend on the candidate: it is computed ONCE per prompt (`set_text_embeds` takes it ready-made;
`encode_text` runs the text tower with torch fp32 matmuls on the GPU from token ids -- the tokenizer's vocabulary files
and the openai/clip-vit-large-patch14 checkpoint are unreachable offline, so both ids and weights are injectable)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from ._lib import ACT_DTYPE
from .ops import Plan
from .scorers import Scorer

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2
TOKEN_ALIGN = 128            # rows per image are padded to a multiple of this (attention K tiles, GEMM M tiles)


def _bicubic(x: float) -> float:
    a = -0.5
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's Resample.c precompute_coeffs + normalize_coeffs_8bpc (bicubic, support 2, box = the whole axis), in the same
    double arithmetic: bounds int32 [out, 2] = (first input index, taps), coefficients int32 [out, ksize] (x 2^22)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_geometry(H: int, W: int, S: int) -> Tuple[int, int, int, int]:
    """CLIPImageProcessor: shortest edge -> S keeping the aspect ratio (int(S * long / short)), centre crop S x S.
    Returns (resized height, resized width, crop top, crop left)."""
    short, long_ = (W, H) if W <= H else (H, W)
    new_long = int(S * long_ / short)
    rh, rw = (new_long, S) if W <= H else (S, new_long)
    return rh, rw, (rh - S) // 2, (rw - S) // 2


def normalise_lut() -> np.ndarray:
    """fp32 [3,256]: rescale + normalise of every uint8 value in the processor's arithmetic (image_transforms.py:
    float32(float64(v) * (1/255)), then (x - mean) / std in float32)."""
    v = (np.arange(256, dtype=np.float64) * 0.00392156862745098).astype(np.float32)
    mean, std = np.array(CLIP_MEAN, dtype=np.float32), np.array(CLIP_STD, dtype=np.float32)
    return ((v[None, :] - mean[:, None]) / std[:, None]).astype(np.float32)


def clip_config_from_state_dict(sd: Dict[str, torch.Tensor]) -> dict:
    pw = sd['vision_model.embeddings.patch_embedding.weight']
    pos = sd['vision_model.embeddings.position_embedding.weight']
    layers = 0
    while f'vision_model.encoder.layers.{layers}.layer_norm1.weight' in sd:
        layers += 1
    G = int(round(math.sqrt(pos.shape[0] - 1)))
    return dict(hidden=pw.shape[0], patch=pw.shape[-1], image_size=G * pw.shape[-1], layers=layers,
                intermediate=sd['vision_model.encoder.layers.0.mlp.fc1.weight'].shape[0],
                proj=sd['visual_projection.weight'].shape[0])


class CLIPVisionPlan:
    """Buffers + kernel plan of the vision tower for a fixed (batch, input height, input width)."""

    def __init__(self, eng: 'CLIPVisionEngine', B: int, H: int, W: int):
        dev, cfg, w = eng.device, eng.cfg, eng.w
        C, S, P, Lp, Kp = cfg['hidden'], cfg['image_size'], cfg['patch'], eng.Lp, eng.Kp
        heads, inter, D = eng.heads, cfg['intermediate'], cfg['proj']
        act = dict(device=dev, dtype=ACT_DTYPE)
        f32 = dict(device=dev, dtype=torch.float32)
        self.B = B
        self.images = torch.zeros(B, 3, H, W, device=dev, dtype=torch.uint8)
        rh, rw, top, left = resize_geometry(H, W, S)
        hb, hk = resample_coeffs(W, rw)
        vb, vk = resample_coeffs(H, rh)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.hb, self.hk = i32(hb[left:left + S]), i32(hk[left:left + S])
        self.vb, self.vk = i32(vb[top:top + S]), i32(vk[top:top + S])
        self.tmp = torch.empty(B, 3, H, S, device=dev, dtype=torch.uint8)
        self.patches = torch.zeros(B * Lp, Kp, **act)          # class-token and padding rows stay zero
        self.text = torch.zeros(1, D, **f32)
        self.pooled = torch.empty(B, C, **f32)
        self.embeds = torch.empty(B, D, **f32)
        self.scores = torch.empty(B, **f32)
        tok = lambda n: torch.empty(B, 1, Lp, n, **act)
        h, h2, n_, qkv, att, f = tok(C), tok(C), tok(C), tok(3 * C), tok(C), tok(inter)
        pos = eng.pos_rows.unsqueeze(0).expand(B, Lp, C).contiguous()   # class + position embeddings as the GEMM's residual
        self.plan = P_ = Plan()
        P_.add_clip_preprocess(self.images, self.tmp, self.patches, self.hb, self.hk, self.vb, self.vk, eng.lut, S, P, Lp)
        P_.add_gemm([self.patches.view(B, 1, Lp, Kp)], [(0, 1, 0, Kp // 64)], w['patch.w'], C, h2, residual=pos,
                    label='patch_embedding', alg_k=3 * P * P)
        P_.add_layernorm(h2, w['pre_ln.weight'], w['pre_ln.bias'], h, label='pre_layernorm')
        self.taps = {'embeddings': h}
        for i in range(cfg['layers']):
            p = f'layers.{i}'
            P_.add_layernorm(h, w[f'{p}.ln1.weight'], w[f'{p}.ln1.bias'], n_, label=f'{p}.ln1')
            P_.add_gemm([n_], [(0, 1, 0, C // 64)], w[f'{p}.qkv.w'], 3 * C, qkv, bias=w[f'{p}.qkv.b'], label=f'{p}.qkv')
            q2 = qkv.view(B * Lp, 3 * C)
            # keys / values through the cross-attention path (a [K | V] column window of the fused projection) for its
            # kv_len masking of the padded token rows; k_col0 / v_col0 are relative to that window
            P_.add_attention(q2, 0, None, att.view(B * Lp, C), B, heads, Lp, v_col0=C, head_dim=64, scale=0.125,
                             kv=q2[:, C:], kv_rows=Lp, kv_len=eng.tokens, kv_div=1, kv_ld=3 * C, label=f'{p}.attn')
            P_.add_gemm([att], [(0, 1, 0, C // 64)], w[f'{p}.out.w'], C, h2, bias=w[f'{p}.out.b'], residual=h, label=f'{p}.out_proj')
            P_.add_layernorm(h2, w[f'{p}.ln2.weight'], w[f'{p}.ln2.bias'], n_, label=f'{p}.ln2')
            P_.add_gemm([n_], [(0, 1, 0, C // 64)], w[f'{p}.fc1.w'], inter, f, bias=w[f'{p}.fc1.b'], act=1, label=f'{p}.fc1+quick_gelu')
            P_.add_gemm([f], [(0, 1, 0, inter // 64)], w[f'{p}.fc2.w'], C, h, bias=w[f'{p}.fc2.b'], residual=h2, label=f'{p}.fc2')
        self.hidden = h
        P_.add_clip_pool_ln(h, Lp * C, w['post_ln.weight'], w['post_ln.bias'], self.pooled, B)
        P_.add_linear(self.pooled, w['proj.weight'], self.embeds, label='visual_projection')
        P_.add_clip_cosine(self.embeds, self.text, self.scores)
        self._keep = (pos, h2, n_, qkv, att, f)
        if eng.use_graphs:
            torch.cuda.synchronize(dev)
            P_.instantiate_graph()


class CLIPVisionEngine:
    """transformers.CLIPModel's vision half (`vision_model.*`, `visual_projection.weight`) on the B200 plan ops."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', use_graphs: bool = True):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('CLIPVisionEngine runs on a B200 only (no CPU fallback)')
        self.cfg = cfg = clip_config_from_state_dict(state_dict)
        C = cfg['hidden']
        if C % 64 or cfg['intermediate'] % 64:
            raise NotImplementedError('CLIP widths must be multiples of 64')
        self.heads = C // 64                                   # every released CLIP ViT uses 64-wide heads
        G = cfg['image_size'] // cfg['patch']
        self.tokens = G * G + 1
        self.Lp = TOKEN_ALIGN * ((self.tokens + TOKEN_ALIGN - 1) // TOKEN_ALIGN)
        self.Kp = 64 * ((3 * cfg['patch'] ** 2 + 63) // 64)
        self.use_graphs = use_graphs
        self.lut = torch.from_numpy(normalise_lut()).to(self.device)
        self._plans: Dict[tuple, CLIPVisionPlan] = {}
        self._pack(state_dict)

    def _pack(self, sd):
        dev, cfg, w = self.device, self.cfg, {}
        C = cfg['hidden']
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        a = lambda t: t.detach().float().contiguous().to(ACT_DTYPE).to(dev)
        pw = sd['vision_model.embeddings.patch_embedding.weight'].detach().float().reshape(C, -1)
        wp = torch.zeros(C, self.Kp)
        wp[:, :pw.shape[1]] = pw
        w['patch.w'] = a(wp)
        pos = torch.zeros(self.Lp, C)
        pos[:self.tokens] = sd['vision_model.embeddings.position_embedding.weight'].detach().float()
        pos[0] += sd['vision_model.embeddings.class_embedding'].detach().float()
        self.pos_rows = a(pos)
        w['pre_ln.weight'], w['pre_ln.bias'] = f(sd['vision_model.pre_layrnorm.weight']), f(sd['vision_model.pre_layrnorm.bias'])
        for i in range(cfg['layers']):
            s, p = f'vision_model.encoder.layers.{i}', f'layers.{i}'
            w[f'{p}.qkv.w'] = a(torch.cat([sd[f'{s}.self_attn.{n}_proj.weight'].detach().float() for n in 'qkv']))
            w[f'{p}.qkv.b'] = f(torch.cat([sd[f'{s}.self_attn.{n}_proj.bias'].detach().float() for n in 'qkv']))
            w[f'{p}.out.w'], w[f'{p}.out.b'] = a(sd[f'{s}.self_attn.out_proj.weight']), f(sd[f'{s}.self_attn.out_proj.bias'])
            for n, m in (('ln1', 'layer_norm1'), ('ln2', 'layer_norm2')):
                w[f'{p}.{n}.weight'], w[f'{p}.{n}.bias'] = f(sd[f'{s}.{m}.weight']), f(sd[f'{s}.{m}.bias'])
            w[f'{p}.fc1.w'], w[f'{p}.fc1.b'] = a(sd[f'{s}.mlp.fc1.weight']), f(sd[f'{s}.mlp.fc1.bias'])
            w[f'{p}.fc2.w'], w[f'{p}.fc2.b'] = a(sd[f'{s}.mlp.fc2.weight']), f(sd[f'{s}.mlp.fc2.bias'])
        w['post_ln.weight'], w['post_ln.bias'] = f(sd['vision_model.post_layernorm.weight']), f(sd['vision_model.post_layernorm.bias'])
        w['proj.weight'] = f(sd['visual_projection.weight'])
        self.w = w

    def plan(self, B: int, H: int, W: int) -> CLIPVisionPlan:
        key = (B, H, W)
        if key not in self._plans:
            self._plans[key] = CLIPVisionPlan(self, B, H, W)
        return self._plans[key]

    @torch.no_grad()
    def run(self, images_u8: torch.Tensor, text_embeds: torch.Tensor) -> CLIPVisionPlan:
        B, _, H, W = images_u8.shape
        cp = self.plan(B, H, W)
        cp.images.copy_(images_u8)
        if text_embeds.shape[0] not in (1, B):
            raise ValueError('one text embedding for all images, or one per image')
        if text_embeds.shape[0] != cp.text.shape[0]:
            raise NotImplementedError('per-image prompts: score each prompt group separately')
        cp.text.copy_(text_embeds)
        cp.plan.run()
        return cp


@torch.no_grad()
def encode_text(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, heads: int, device='cuda') -> torch.Tensor:
    """`CLIPModel.get_text_features` (modeling_clip.py CLIPTextTransformer + text_projection) for EOS-terminated, unpadded
    token ids [1, T]: causal attention, pooled at argmax(ids).  ONCE per prompt -- not part of the per-candidate path --
    so it is plain torch fp32 on the GPU."""
    dev = torch.device(device)
    g = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32)
    ids = input_ids.to(dev)
    T = ids.shape[1]
    h = g('text_model.embeddings.token_embedding.weight')[ids] + g('text_model.embeddings.position_embedding.weight')[:T]
    C = h.shape[-1]
    hd = C // heads
    mask = torch.full((T, T), float('-inf'), device=dev).triu(1)
    i = 0
    while f'text_model.encoder.layers.{i}.layer_norm1.weight' in sd:
        p = f'text_model.encoder.layers.{i}'
        x = F.layer_norm(h, (C,), g(f'{p}.layer_norm1.weight'), g(f'{p}.layer_norm1.bias'), 1e-5)
        q, k, v = (F.linear(x, g(f'{p}.self_attn.{n}_proj.weight'), g(f'{p}.self_attn.{n}_proj.bias')).view(1, T, heads, hd).transpose(1, 2)
                   for n in 'qkv')
        a = (torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2) + mask, dim=-1) @ v).transpose(1, 2).reshape(1, T, C)
        h = h + F.linear(a, g(f'{p}.self_attn.out_proj.weight'), g(f'{p}.self_attn.out_proj.bias'))
        x = F.layer_norm(h, (C,), g(f'{p}.layer_norm2.weight'), g(f'{p}.layer_norm2.bias'), 1e-5)
        x = F.linear(x, g(f'{p}.mlp.fc1.weight'), g(f'{p}.mlp.fc1.bias'))
        h = h + F.linear(x * torch.sigmoid(1.702 * x), g(f'{p}.mlp.fc2.weight'), g(f'{p}.mlp.fc2.bias'))
        i += 1
    h = F.layer_norm(h, (C,), g('text_model.final_layer_norm.weight'), g('text_model.final_layer_norm.bias'), 1e-5)
    return F.linear(h[torch.arange(1), ids.argmax(dim=-1)], g('text_projection.weight'))


class CLIPScorer(Scorer):
    """Same call protocol as the reference's CLIPScorer (sd/scorers.py:149-213): `scorer(images uint8 [M,3,H,W], prompts,
    timesteps) -> cosine similarity [M]`.  `state_dict` = transformers.CLIPModel's (the reference downloads
    openai/clip-vit-large-patch14, unreachable offline).  The prompt embedding comes from `tokenize` (a callable
    prompt -> token ids [1,T], e.g. a CLIPTokenizer loaded from local files) + the text tower, or is given directly with
    `set_text_embeds(prompt, embeds [1, D])`; it is cached per prompt."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, dtype=torch.float32, device='cuda', tokenize=None,
                 text_heads: int = 12):
        super().__init__(dtype)
        if state_dict is None:
            raise RuntimeError('CLIPScorer needs the CLIPModel state dict (openai/clip-vit-large-patch14 cannot be downloaded '
                               'offline): pass state_dict=CLIPModel.from_pretrained(path).state_dict()')
        self.device = torch.device(device)
        self.engine = CLIPVisionEngine(state_dict, device=device)
        self._text_sd = {k: v for k, v in state_dict.items() if k.startswith('text_') or k.startswith('text_projection')}
        self.tokenize, self.text_heads = tokenize, text_heads
        self._text: Dict[str, torch.Tensor] = {}

    def set_text_embeds(self, prompt: str, embeds: torch.Tensor):
        self._text[prompt] = embeds.detach().to(device=self.device, dtype=torch.float32).reshape(1, -1).contiguous()

    def text_embeds(self, prompt: str) -> torch.Tensor:
        if prompt not in self._text:
            if self.tokenize is None:
                raise RuntimeError('CLIPScorer: no embedding for this prompt -- call set_text_embeds(prompt, embeds) or pass '
                                   'tokenize= (the CLIP vocabulary files are not available offline)')
            self.set_text_embeds(prompt, encode_text(self._text_sd, self.tokenize(prompt), self.text_heads, self.device))
        return self._text[prompt]

    @torch.no_grad()
    def __call__(self, images, prompts, timesteps=None):
        if not isinstance(images, torch.Tensor) or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[1] != 3:
            raise TypeError('B200 CLIPScorer scores uint8 [M,3,H,W] images (pipeline_stable_diffusion.py:1115)')
        images = images.to(self.device).contiguous()
        M = images.shape[0]
        if prompts is None:                                    # sd/scorers.py:181-183
            return torch.zeros(M, device=self.device, dtype=self.dtype)
        if not isinstance(prompts, (list, tuple)):
            prompts = [prompts] * M
        elif len(prompts) == 1 and M > 1:
            prompts = list(prompts) * M
        if len(prompts) != M:
            raise ValueError('one prompt per image, or one for all')
        out = torch.empty(M, device=self.device, dtype=torch.float32)
        groups: Dict[str, List[int]] = {}
        for i, p in enumerate(prompts):
            groups.setdefault(p, []).append(i)
        for p, idx in groups.items():                          # the search passes ONE prompt for all candidates: one group
            sel = images if len(idx) == M else images[torch.tensor(idx, device=self.device)]
            cp = self.engine.run(sel, self.text_embeds(p))
            if len(idx) == M:
                out.copy_(cp.scores)
            else:
                out[torch.tensor(idx, device=self.device)] = cp.scores
        return out.to(self.dtype)

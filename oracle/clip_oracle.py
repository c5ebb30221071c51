"""CPU restatement of the reference's CLIPScorer (sd/scorers.py:149-213) -- TEST INFRASTRUCTURE ONLY.

The scorer's arithmetic lives in third-party code that is NOT under /root/reference: HuggingFace `transformers`
(`CLIPModel`, `CLIPProcessor`; the reference pins no version, this image has 5.5.0) and Pillow's resampler (the slow /
PIL image processor resizes with `PIL.Image.resize(..., BICUBIC)`).  Restated here from their published algorithms:

  * `pil_bicubic_resize_u8`   Pillow's two-pass 8-bit resampler (src/libImaging/Resample.c: precompute_coeffs,
                              normalize_coeffs_8bpc with PRECISION_BITS = 22, ImagingResampleHorizontal/Vertical_8bpc) in
                              integer arithmetic: bit-exact against PIL (tests/test_clip_oracle.py);
  * `clip_preprocess`         CLIPImageProcessor: resize shortest edge -> center crop -> x/255 -> (x - mean) / std;
  * `clip_vision_forward`     CLIPVisionTransformer + visual_projection (modeling_clip.py): patch conv, class token,
                              position embedding, pre-LN, N x [LN, MHA, +res, LN, fc1, quick_gelu, fc2, +res], post-LN of
                              the class token, projection;
  * `clip_text_forward`       CLIPTextTransformer + text_projection (causal mask, EOS pooling = argmax of the ids);
  * `clip_score`              cosine similarity of the normalised embeddings (sd/scorers.py:176-213).
Pinned against `transformers.CLIPModel` / `CLIPImageProcessorPil` themselves run in this container on seeded weights
(oracle/make_golden_clip.py -> tests/golden/clip_*.pt)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def clip_param_shapes(hidden=1024, layers=24, heads=16, intermediate=4096, image_size=224, patch=14, proj=768,
                      t_hidden=768, t_layers=12, t_intermediate=3072, vocab=49408, max_pos=77) -> Dict[str, Tuple[int, ...]]:
    """State-dict names / shapes of transformers.CLIPModel (defaults: openai/clip-vit-large-patch14)."""
    shp: Dict[str, Tuple[int, ...]] = {'logit_scale': ()}
    P = (image_size // patch) ** 2

    def tower(prefix, h, inter, n):
        for i in range(n):
            p = f'{prefix}.encoder.layers.{i}'
            for nm in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
                shp[f'{p}.self_attn.{nm}.weight'], shp[f'{p}.self_attn.{nm}.bias'] = (h, h), (h,)
            for nm in ('layer_norm1', 'layer_norm2'):
                shp[f'{p}.{nm}.weight'] = shp[f'{p}.{nm}.bias'] = (h,)
            shp[f'{p}.mlp.fc1.weight'], shp[f'{p}.mlp.fc1.bias'] = (inter, h), (inter,)
            shp[f'{p}.mlp.fc2.weight'], shp[f'{p}.mlp.fc2.bias'] = (h, inter), (h,)

    shp['vision_model.embeddings.class_embedding'] = (hidden,)
    shp['vision_model.embeddings.patch_embedding.weight'] = (hidden, 3, patch, patch)
    shp['vision_model.embeddings.position_embedding.weight'] = (P + 1, hidden)
    shp['vision_model.pre_layrnorm.weight'] = shp['vision_model.pre_layrnorm.bias'] = (hidden,)
    tower('vision_model', hidden, intermediate, layers)
    shp['vision_model.post_layernorm.weight'] = shp['vision_model.post_layernorm.bias'] = (hidden,)
    shp['visual_projection.weight'] = (proj, hidden)
    shp['text_model.embeddings.token_embedding.weight'] = (vocab, t_hidden)
    shp['text_model.embeddings.position_embedding.weight'] = (max_pos, t_hidden)
    tower('text_model', t_hidden, t_intermediate, t_layers)
    shp['text_model.final_layer_norm.weight'] = shp['text_model.final_layer_norm.bias'] = (t_hidden,)
    shp['text_projection.weight'] = (proj, t_hidden)
    return shp


def seeded_clip_state_dict(shapes, seed: int) -> Dict[str, torch.Tensor]:
    """N(0, 1/fan_in) matrices, N(0, 0.02) embeddings, LayerNorm gains 1 + N(0, 0.1^2), biases N(0, 0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        s = shapes[name]
        if len(s) == 0:
            out[name] = torch.tensor(2.6592)
        elif 'embedding' in name:
            out[name] = torch.randn(s, generator=g) * (0.02 if 'patch' not in name else 1.0 / math.sqrt(int(np.prod(s[1:]))))
        elif len(s) == 1:
            t = torch.randn(s, generator=g) * 0.1
            out[name] = t + 1.0 if name.endswith('weight') else t
        else:
            out[name] = torch.randn(s, generator=g) / math.sqrt(s[1])
    return out


# ---------------------------------------------------------------------------------------------- Pillow's resampler
def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resize_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2), box = the whole axis.
    Returns (bounds int32 [out, 2] = (xmin, count), coefficients int32 [out, ksize])."""
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
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis_u8(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One pass of ImagingResample{Horizontal,Vertical}_8bpc along `axis` of a uint8 array."""
    bounds, kk = pil_resize_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (src[..., xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def pil_bicubic_resize_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H, W, C] -> [out_h, out_w, C] exactly as PIL.Image.resize((out_w, out_h), BICUBIC): horizontal pass
    first, then vertical, each rounded to uint8 (ImagingResample, Resample.c)."""
    if img.shape[1] != out_w:
        img = _resample_axis_u8(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _resample_axis_u8(img, out_h, 0)
    return img


def clip_preprocess(images: torch.Tensor, size: int = 224) -> torch.Tensor:
    """CLIPImageProcessor on uint8 [B, 3, H, W]: shortest edge -> `size` (bicubic, PIL), center crop size x size, /255,
    (x - mean) / std.  Returns fp32 [B, 3, size, size]."""
    out = []
    for im in images:
        a = im.permute(1, 2, 0).numpy()
        h, w = a.shape[:2]
        short, long_ = (w, h) if w <= h else (h, w)
        new_short, new_long = size, int(size * long_ / short)
        nh, nw = (new_long, new_short) if w <= h else (new_short, new_long)
        a = pil_bicubic_resize_u8(a, nh, nw)
        top, left = (nh - size) // 2, (nw - size) // 2
        a = a[top:top + size, left:left + size]
        out.append(torch.from_numpy(normalise_lut()[np.arange(3)[:, None, None], a.transpose(2, 0, 1)]))
    return torch.stack(out)


def normalise_lut() -> np.ndarray:
    """fp32 [3, 256]: the processor's rescale + normalise of every possible uint8 value, in ITS arithmetic
    (image_transforms.py rescale: float32(float64(v) * (1/255)); normalize: (x - mean) / std in float32)."""
    v = (np.arange(256, dtype=np.float64) * 0.00392156862745098).astype(np.float32)
    mean, std = np.array(CLIP_MEAN, dtype=np.float32), np.array(CLIP_STD, dtype=np.float32)
    return ((v[None, :] - mean[:, None]) / std[:, None]).astype(np.float32)


# ---------------------------------------------------------------------------------------------- towers
def _encoder(sd, prefix: str, h: torch.Tensor, heads: int, mask=None) -> torch.Tensor:
    i = 0
    while f'{prefix}.encoder.layers.{i}.layer_norm1.weight' in sd:
        p = f'{prefix}.encoder.layers.{i}'
        C = h.shape[-1]
        hd = C // heads
        r = h
        x = F.layer_norm(h, (C,), sd[f'{p}.layer_norm1.weight'], sd[f'{p}.layer_norm1.bias'], 1e-5)
        q = F.linear(x, sd[f'{p}.self_attn.q_proj.weight'], sd[f'{p}.self_attn.q_proj.bias']) * hd ** -0.5
        k = F.linear(x, sd[f'{p}.self_attn.k_proj.weight'], sd[f'{p}.self_attn.k_proj.bias'])
        v = F.linear(x, sd[f'{p}.self_attn.v_proj.weight'], sd[f'{p}.self_attn.v_proj.bias'])
        B, L, _ = x.shape
        q, k, v = (t.view(B, L, heads, hd).transpose(1, 2) for t in (q, k, v))
        w = q @ k.transpose(-1, -2)
        if mask is not None:
            w = w + mask
        a = (torch.softmax(w, dim=-1) @ v).transpose(1, 2).reshape(B, L, C)
        h = r + F.linear(a, sd[f'{p}.self_attn.out_proj.weight'], sd[f'{p}.self_attn.out_proj.bias'])
        r = h
        x = F.layer_norm(h, (C,), sd[f'{p}.layer_norm2.weight'], sd[f'{p}.layer_norm2.bias'], 1e-5)
        x = F.linear(x, sd[f'{p}.mlp.fc1.weight'], sd[f'{p}.mlp.fc1.bias'])
        x = x * torch.sigmoid(1.702 * x)                                   # quick_gelu (activations.py QuickGELUActivation)
        h = r + F.linear(x, sd[f'{p}.mlp.fc2.weight'], sd[f'{p}.mlp.fc2.bias'])
        i += 1
    return h


def clip_vision_forward(sd, pixel_values: torch.Tensor, heads: int) -> torch.Tensor:
    """CLIPModel.get_image_features (modeling_clip.py): fp32 [B, 3, S, S] -> [B, proj]."""
    w = sd['vision_model.embeddings.patch_embedding.weight']
    x = F.conv2d(pixel_values, w, stride=w.shape[-1]).flatten(2).transpose(1, 2)              # [B, P, C]
    cls = sd['vision_model.embeddings.class_embedding'].expand(x.shape[0], 1, -1)
    h = torch.cat([cls, x], dim=1) + sd['vision_model.embeddings.position_embedding.weight'].unsqueeze(0)
    C = h.shape[-1]
    h = F.layer_norm(h, (C,), sd['vision_model.pre_layrnorm.weight'], sd['vision_model.pre_layrnorm.bias'], 1e-5)
    h = _encoder(sd, 'vision_model', h, heads)
    pooled = F.layer_norm(h[:, 0], (C,), sd['vision_model.post_layernorm.weight'], sd['vision_model.post_layernorm.bias'], 1e-5)
    return F.linear(pooled, sd['visual_projection.weight'])


def clip_text_forward(sd, input_ids: torch.Tensor, heads: int) -> torch.Tensor:
    """CLIPModel.get_text_features for unpadded / EOS-terminated ids (causal mask; pooled at argmax(ids), the EOS token of
    the original vocabulary): int64 [B, T] -> [B, proj]."""
    T = input_ids.shape[1]
    h = sd['text_model.embeddings.token_embedding.weight'][input_ids] + sd['text_model.embeddings.position_embedding.weight'][:T]
    mask = torch.full((T, T), float('-inf')).triu(1)
    h = _encoder(sd, 'text_model', h, heads, mask=mask)
    C = h.shape[-1]
    h = F.layer_norm(h, (C,), sd['text_model.final_layer_norm.weight'], sd['text_model.final_layer_norm.bias'], 1e-5)
    pooled = h[torch.arange(h.shape[0]), input_ids.argmax(dim=-1)]
    return F.linear(pooled, sd['text_projection.weight'])


def clip_score(image_embeds: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
    """sd/scorers.py:176-213: cosine similarity of the L2-normalised embeddings, one prompt per image (or one for all)."""
    i = image_embeds / torch.linalg.vector_norm(image_embeds, dim=-1, keepdim=True)
    t = text_embeds / torch.linalg.vector_norm(text_embeds, dim=-1, keepdim=True)
    if t.shape[0] == 1 and i.shape[0] > 1:
        t = t.expand(i.shape[0], -1)
    return torch.sum(i * t, dim=1)

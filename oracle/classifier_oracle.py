"""CPU oracle for the ImageNet classifier scorer.  TEST INFRASTRUCTURE ONLY.

Functional restatement of EncoderUNetModel (edm/unet.py:701-912) with the configuration
ImageNetScorer fixes (edm/scorers.py:77-86,127-140): scale-shift ResBlocks (edm/unet.py:254-274),
ResBlock down-sampling by 2x2 average pool (:211-213), AttentionBlock + QKVAttentionLegacy
(:317-323, :355-372), GroupNorm32 (edm/nn_utils.py:17-19), timestep_embedding (:103-121) and the
attention pool (edm/unet.py:40-69, :388-407); plus ImageNetScorer.__call__ (edm/scorers.py:143-174).
Pinned against the reference by tests/golden/classifier_*.pt (oracle/make_golden.py)."""
import math
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F


def classifier_layout(image_size=64, model_channels=128, num_res_blocks=4, attention_resolutions=(2, 4, 8),
                      channel_mult=(1, 2, 3, 4), **_unused):
    """[(prefix, cin, cout, down, attention)] for input_blocks[1:] and middle_block, in execution order."""
    blocks = []
    ch = int(channel_mult[0] * model_channels)
    ds, idx = 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            cout = int(mult * model_channels)
            blocks.append((f'input_blocks.{idx}', ch, cout, False, ds in attention_resolutions))
            ch = cout
            idx += 1
        if level != len(channel_mult) - 1:
            blocks.append((f'input_blocks.{idx}', ch, ch, True, False))
            idx += 1
            ds *= 2
    return blocks, ch, image_size // ds


def classifier_param_shapes(image_size=64, in_channels=3, model_channels=128, out_channels=1000, num_res_blocks=4,
                            attention_resolutions=(2, 4, 8), channel_mult=(1, 2, 3, 4), **_unused) -> Dict[str, Tuple[int, ...]]:
    E = model_channels * 4
    shp = {'time_embed.0.weight': (E, model_channels), 'time_embed.0.bias': (E,), 'time_embed.2.weight': (E, E),
           'time_embed.2.bias': (E,)}
    c0 = int(channel_mult[0] * model_channels)
    shp['input_blocks.0.0.weight'], shp['input_blocks.0.0.bias'] = (c0, in_channels, 3, 3), (c0,)

    def res(p, cin, cout):
        shp[f'{p}.in_layers.0.weight'] = shp[f'{p}.in_layers.0.bias'] = (cin,)
        shp[f'{p}.in_layers.2.weight'], shp[f'{p}.in_layers.2.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{p}.emb_layers.1.weight'], shp[f'{p}.emb_layers.1.bias'] = (2 * cout, E), (2 * cout,)
        shp[f'{p}.out_layers.0.weight'] = shp[f'{p}.out_layers.0.bias'] = (cout,)
        shp[f'{p}.out_layers.3.weight'], shp[f'{p}.out_layers.3.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shp[f'{p}.skip_connection.weight'], shp[f'{p}.skip_connection.bias'] = (cout, cin, 1, 1), (cout,)

    def attn(p, c):
        shp[f'{p}.norm.weight'] = shp[f'{p}.norm.bias'] = (c,)
        shp[f'{p}.qkv.weight'], shp[f'{p}.qkv.bias'] = (3 * c, c, 1), (3 * c,)
        shp[f'{p}.proj_out.weight'], shp[f'{p}.proj_out.bias'] = (c, c, 1), (c,)

    blocks, ch, sp = classifier_layout(image_size, model_channels, num_res_blocks, attention_resolutions, channel_mult)
    for p, cin, cout, down, at in blocks:
        res(f'{p}.0', cin, cout)
        if at:
            attn(f'{p}.1', cout)
    res('middle_block.0', ch, ch)
    attn('middle_block.1', ch)
    res('middle_block.2', ch, ch)
    shp['out.0.weight'] = shp['out.0.bias'] = (ch,)
    shp['out.2.positional_embedding'] = (ch, sp * sp + 1)
    shp['out.2.qkv_proj.weight'], shp['out.2.qkv_proj.bias'] = (3 * ch, ch, 1), (3 * ch,)
    shp['out.2.c_proj.weight'], shp['out.2.c_proj.bias'] = (out_channels, ch, 1), (out_channels,)
    return shp


def seeded_classifier_state_dict(shapes, seed):
    """Same convention as edm_oracle.seeded_state_dict (every tensor non-degenerate; the reference's
    zero_module convs, edm/unet.py:228,312, would otherwise switch half the network off)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 1:
            t = torch.randn(shp, generator=g) * 0.1
            out[name] = t + 1.0 if name.endswith('weight') else t
        elif name.endswith('positional_embedding'):
            out[name] = torch.randn(shp, generator=g) / math.sqrt(shp[0])
        else:
            out[name] = torch.randn(shp, generator=g) / math.sqrt(int(math.prod(shp[1:])))
    return out


def timestep_embedding(timesteps, dim, max_period=10000):
    """edm/nn_utils.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn32(sd, p, x):
    return F.group_norm(x.float(), 32, sd[f'{p}.weight'], sd[f'{p}.bias'], 1e-5)


def _res_block(sd, p, x, emb, down):
    """ResBlock._forward, edm/unet.py:254-274 (use_scale_shift_norm=True)."""
    h = F.silu(_gn32(sd, f'{p}.in_layers.0', x))
    if down:
        h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    h = F.conv2d(h, sd[f'{p}.in_layers.2.weight'], sd[f'{p}.in_layers.2.bias'], padding=1)
    e = F.linear(F.silu(emb), sd[f'{p}.emb_layers.1.weight'], sd[f'{p}.emb_layers.1.bias'])[:, :, None, None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = _gn32(sd, f'{p}.out_layers.0', h) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[f'{p}.out_layers.3.weight'], sd[f'{p}.out_layers.3.bias'], padding=1)
    if f'{p}.skip_connection.weight' in sd:
        x = F.conv2d(x, sd[f'{p}.skip_connection.weight'], sd[f'{p}.skip_connection.bias'])
    return x + h


def _attn_block(sd, p, x, head_channels=64):
    """AttentionBlock._forward + QKVAttentionLegacy, edm/unet.py:317-323, 355-372."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn32(sd, f'{p}.norm', xf), sd[f'{p}.qkv.weight'], sd[f'{p}.qkv.bias'])
    heads = c // head_channels
    ch = head_channels
    q, k, v = qkv.reshape(b * heads, ch * 3, -1).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum('bct,bcs->bts', q * scale, k * scale).float(), dim=-1)
    a = torch.einsum('bts,bcs->bct', w, v).reshape(b, -1, xf.shape[-1])
    h = F.conv1d(a, sd[f'{p}.proj_out.weight'], sd[f'{p}.proj_out.bias'])
    return (xf + h).reshape(b, c, hh, ww)


def _attention_pool(sd, p, x, head_channels=64):
    """AttentionPool2d.forward + QKVAttention, edm/unet.py:61-69, 388-407."""
    b, c = x.shape[:2]
    x = x.reshape(b, c, -1)
    x = torch.cat([x.mean(dim=-1, keepdim=True), x], dim=-1) + sd[f'{p}.positional_embedding'][None]
    qkv = F.conv1d(x, sd[f'{p}.qkv_proj.weight'], sd[f'{p}.qkv_proj.bias'])
    heads = c // head_channels
    ch = head_channels
    length = qkv.shape[-1]
    q, k, v = qkv.chunk(3, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum('bct,bcs->bts', (q * scale).reshape(b * heads, ch, length),
                                   (k * scale).reshape(b * heads, ch, length)).float(), dim=-1)
    a = torch.einsum('bts,bcs->bct', w, v.reshape(b * heads, ch, length)).reshape(b, -1, length)
    return F.conv1d(a, sd[f'{p}.c_proj.weight'], sd[f'{p}.c_proj.bias'])[:, :, 0]


@torch.no_grad()
def classifier_logits(sd, cfg, x, timesteps):
    """EncoderUNetModel.forward, edm/unet.py:889-912 (pool='attention')."""
    emb = timestep_embedding(timesteps, cfg.get('model_channels', 128))
    emb = F.linear(F.silu(F.linear(emb, sd['time_embed.0.weight'], sd['time_embed.0.bias'])),
                   sd['time_embed.2.weight'], sd['time_embed.2.bias'])
    h = F.conv2d(x.float(), sd['input_blocks.0.0.weight'], sd['input_blocks.0.0.bias'], padding=1)
    blocks, ch, sp = classifier_layout(**cfg)
    for p, cin, cout, down, at in blocks:
        h = _res_block(sd, f'{p}.0', h, emb, down)
        if at:
            h = _attn_block(sd, f'{p}.1', h)
    h = _res_block(sd, 'middle_block.0', h, emb, False)
    h = _attn_block(sd, 'middle_block.1', h)
    h = _res_block(sd, 'middle_block.2', h, emb, False)
    return _attention_pool(sd, 'out.2', F.silu(_gn32(sd, 'out.0', h)))


@torch.no_grad()
def imagenet_score(sd, cfg, images, class_labels, timesteps):
    """ImageNetScorer.__call__, edm/scorers.py:143-174: softmax PROBABILITY of the target class."""
    if images.dtype == torch.uint8:
        images = images.float() / 255.0
    probs = F.softmax(classifier_logits(sd, cfg, images, timesteps), dim=1)
    target = torch.argmax(class_labels, dim=1) if class_labels.dim() > 1 else class_labels
    return probs[torch.arange(probs.size(0)), target]

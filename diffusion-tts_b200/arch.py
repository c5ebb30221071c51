"""Parameter inventory of the ADM (DhariwalUNet) architecture preset the benchmark and the CLI
use when no pretrained pickle is reachable: names and shapes as the reference registers them
(edm/training/networks.py:373-433; preset edm/train.py:124), plus a seeded random init."""
import math
from typing import Dict, Sequence, Tuple

import torch


def adm_param_shapes(img_resolution=64, in_channels=3, out_channels=3, label_dim=1000, model_channels=192,
                     channel_mult: Sequence[int] = (1, 2, 3, 4), channel_mult_emb=4, num_blocks=3,
                     attn_resolutions: Sequence[int] = (32, 16, 8)) -> Dict[str, Tuple[int, ...]]:
    E = model_channels * channel_mult_emb
    shp: Dict[str, Tuple[int, ...]] = {
        'map_layer0.weight': (E, model_channels), 'map_layer0.bias': (E,),
        'map_layer1.weight': (E, E), 'map_layer1.bias': (E,),
    }
    if label_dim:
        shp['map_label.weight'] = (E, label_dim)

    def unet_block(name, cin, cout, attention):
        shp[f'{name}.norm0.weight'] = shp[f'{name}.norm0.bias'] = (cin,)
        shp[f'{name}.conv0.weight'], shp[f'{name}.conv0.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{name}.affine.weight'], shp[f'{name}.affine.bias'] = (2 * cout, E), (2 * cout,)
        shp[f'{name}.norm1.weight'] = shp[f'{name}.norm1.bias'] = (cout,)
        shp[f'{name}.conv1.weight'], shp[f'{name}.conv1.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shp[f'{name}.skip.weight'], shp[f'{name}.skip.bias'] = (cout, cin, 1, 1), (cout,)
        if attention:
            shp[f'{name}.norm2.weight'] = shp[f'{name}.norm2.bias'] = (cout,)
            shp[f'{name}.qkv.weight'], shp[f'{name}.qkv.bias'] = (3 * cout, cout, 1, 1), (3 * cout,)
            shp[f'{name}.proj.weight'], shp[f'{name}.proj.bias'] = (cout, cout, 1, 1), (cout,)

    skips, c = [], in_channels
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            c_new = model_channels * mult
            shp[f'enc.{res}x{res}_conv.weight'], shp[f'enc.{res}x{res}_conv.bias'] = (c_new, c, 3, 3), (c_new,)
            c = c_new
        else:
            unet_block(f'enc.{res}x{res}_down', c, c, False)
        skips.append(c)
        for idx in range(num_blocks):
            c_new = model_channels * mult
            unet_block(f'enc.{res}x{res}_block{idx}', c, c_new, res in attn_resolutions)
            c = c_new
            skips.append(c)
    for level, mult in reversed(list(enumerate(channel_mult))):
        res = img_resolution >> level
        if level == len(channel_mult) - 1:
            unet_block(f'dec.{res}x{res}_in0', c, c, True)
            unet_block(f'dec.{res}x{res}_in1', c, c, False)
        else:
            unet_block(f'dec.{res}x{res}_up', c, c, False)
        for idx in range(num_blocks + 1):
            c_new = model_channels * mult
            unet_block(f'dec.{res}x{res}_block{idx}', c + skips.pop(), c_new, res in attn_resolutions)
            c = c_new
    shp['out_norm.weight'] = shp['out_norm.bias'] = (c,)
    shp['out_conv.weight'], shp['out_conv.bias'] = (out_channels, c, 3, 3), (out_channels,)
    return shp


def random_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Every tensor ~ N(0, 1/fan_in) (biases N(0, 0.01), norm gains 1 + N(0, 0.01)) from one seeded
    CPU generator in sorted-name order: the same convention as the parity fixtures."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 1:
            t = torch.randn(shp, generator=g) * 0.1
            sd[name] = t + 1.0 if name.endswith('weight') else t
        else:
            sd[name] = torch.randn(shp, generator=g) / math.sqrt(int(math.prod(shp[1:])))
    return sd

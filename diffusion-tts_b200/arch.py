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


def classifier_param_shapes(image_size=64, in_channels=3, model_channels=128, out_channels=1000, num_res_blocks=4,
                            attention_ds: Sequence[int] = (2, 4, 8), channel_mult: Sequence[int] = (1, 2, 3, 4)):
    """Parameter inventory of the ADM classifier the reference's ImageNetScorer builds
    (EncoderUNetModel kwargs at edm/scorers.py:127-140; module registration edm/unet.py:747-866)."""
    E = model_channels * 4
    shp: Dict[str, Tuple[int, ...]] = {'time_embed.0.weight': (E, model_channels), 'time_embed.0.bias': (E,),
                                       'time_embed.2.weight': (E, E), 'time_embed.2.bias': (E,)}
    ch = int(channel_mult[0] * model_channels)
    shp['input_blocks.0.0.weight'], shp['input_blocks.0.0.bias'] = (ch, in_channels, 3, 3), (ch,)

    def res_block(p, cin, cout):
        shp[f'{p}.in_layers.0.weight'] = shp[f'{p}.in_layers.0.bias'] = (cin,)
        shp[f'{p}.in_layers.2.weight'], shp[f'{p}.in_layers.2.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{p}.emb_layers.1.weight'], shp[f'{p}.emb_layers.1.bias'] = (2 * cout, E), (2 * cout,)
        shp[f'{p}.out_layers.0.weight'] = shp[f'{p}.out_layers.0.bias'] = (cout,)
        shp[f'{p}.out_layers.3.weight'], shp[f'{p}.out_layers.3.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shp[f'{p}.skip_connection.weight'], shp[f'{p}.skip_connection.bias'] = (cout, cin, 1, 1), (cout,)

    def attn_block(p, c):
        shp[f'{p}.norm.weight'] = shp[f'{p}.norm.bias'] = (c,)
        shp[f'{p}.qkv.weight'], shp[f'{p}.qkv.bias'] = (3 * c, c, 1), (3 * c,)
        shp[f'{p}.proj_out.weight'], shp[f'{p}.proj_out.bias'] = (c, c, 1), (c,)

    ds, idx = 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            cout = int(mult * model_channels)
            res_block(f'input_blocks.{idx}.0', ch, cout)
            if ds in attention_ds:
                attn_block(f'input_blocks.{idx}.1', cout)
            ch = cout
            idx += 1
        if level != len(channel_mult) - 1:
            res_block(f'input_blocks.{idx}.0', ch, ch)
            idx += 1
            ds *= 2
    res_block('middle_block.0', ch, ch)
    attn_block('middle_block.1', ch)
    res_block('middle_block.2', ch, ch)
    side = image_size // ds
    shp['out.0.weight'] = shp['out.0.bias'] = (ch,)
    shp['out.2.positional_embedding'] = (ch, side * side + 1)
    shp['out.2.qkv_proj.weight'], shp['out.2.qkv_proj.bias'] = (3 * ch, ch, 1), (3 * ch,)
    shp['out.2.c_proj.weight'], shp['out.2.c_proj.bias'] = (out_channels, ch, 1), (out_channels,)
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


def sd_unet_param_shapes(in_channels=4, out_channels=4, block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
                         layers_per_block=2, cross_attn_down: Sequence[bool] = (True, True, True, False),
                         cross_attention_dim=768) -> Dict[str, Tuple[int, ...]]:
    """Parameter inventory of the SD-1.5-shaped `UNet2DConditionModel` (BASELINE.json config 5): names and shapes as the
    vendored diffusers registers them (sd/diffusers/src/diffusers/models/unets/unet_2d_condition.py:237-480,
    unet_2d_blocks.py CrossAttnDownBlock2D/DownBlock2D/UNetMidBlock2DCrossAttn/UpBlock2D/CrossAttnUpBlock2D,
    resnet.py ResnetBlock2D, transformers/transformer_2d.py, attention.py BasicTransformerBlock).  The up path mirrors
    the down path (`cross_attn_down` reversed).  tests/golden/sd_unet_shapes.json pins this against the reference."""
    boc = list(block_out_channels)
    E = boc[0] * 4
    shp: Dict[str, Tuple[int, ...]] = {
        'conv_in.weight': (boc[0], in_channels, 3, 3), 'conv_in.bias': (boc[0],),
        'time_embedding.linear_1.weight': (E, boc[0]), 'time_embedding.linear_1.bias': (E,),
        'time_embedding.linear_2.weight': (E, E), 'time_embedding.linear_2.bias': (E,),
    }

    def resnet(p, cin, cout):
        shp[f'{p}.norm1.weight'] = shp[f'{p}.norm1.bias'] = (cin,)
        shp[f'{p}.conv1.weight'], shp[f'{p}.conv1.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{p}.time_emb_proj.weight'], shp[f'{p}.time_emb_proj.bias'] = (cout, E), (cout,)
        shp[f'{p}.norm2.weight'] = shp[f'{p}.norm2.bias'] = (cout,)
        shp[f'{p}.conv2.weight'], shp[f'{p}.conv2.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shp[f'{p}.conv_shortcut.weight'], shp[f'{p}.conv_shortcut.bias'] = (cout, cin, 1, 1), (cout,)

    def transformer(p, c):
        shp[f'{p}.norm.weight'] = shp[f'{p}.norm.bias'] = (c,)
        shp[f'{p}.proj_in.weight'], shp[f'{p}.proj_in.bias'] = (c, c, 1, 1), (c,)
        t = f'{p}.transformer_blocks.0'
        for nm in ('norm1', 'norm2', 'norm3'):
            shp[f'{t}.{nm}.weight'] = shp[f'{t}.{nm}.bias'] = (c,)
        for a, kdim in (('attn1', c), ('attn2', cross_attention_dim)):
            shp[f'{t}.{a}.to_q.weight'] = (c, c)
            shp[f'{t}.{a}.to_k.weight'] = shp[f'{t}.{a}.to_v.weight'] = (c, kdim)
            shp[f'{t}.{a}.to_out.0.weight'], shp[f'{t}.{a}.to_out.0.bias'] = (c, c), (c,)
        shp[f'{t}.ff.net.0.proj.weight'], shp[f'{t}.ff.net.0.proj.bias'] = (8 * c, c), (8 * c,)
        shp[f'{t}.ff.net.2.weight'], shp[f'{t}.ff.net.2.bias'] = (c, 4 * c), (c,)
        shp[f'{p}.proj_out.weight'], shp[f'{p}.proj_out.bias'] = (c, c, 1, 1), (c,)

    out_c = boc[0]
    for i, c in enumerate(boc):
        in_c, out_c = out_c, c
        for j in range(layers_per_block):
            resnet(f'down_blocks.{i}.resnets.{j}', in_c if j == 0 else out_c, out_c)
            if cross_attn_down[i]:
                transformer(f'down_blocks.{i}.attentions.{j}', out_c)
        if i != len(boc) - 1:
            shp[f'down_blocks.{i}.downsamplers.0.conv.weight'] = (out_c, out_c, 3, 3)
            shp[f'down_blocks.{i}.downsamplers.0.conv.bias'] = (out_c,)
    resnet('mid_block.resnets.0', boc[-1], boc[-1])
    transformer('mid_block.attentions.0', boc[-1])
    resnet('mid_block.resnets.1', boc[-1], boc[-1])
    rev = boc[::-1]
    cross_up = list(cross_attn_down)[::-1]
    out_c = rev[0]
    for i in range(len(rev)):
        prev_out, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(rev) - 1)]
        for j in range(layers_per_block + 1):
            skip_c = in_c if j == layers_per_block else out_c
            resnet(f'up_blocks.{i}.resnets.{j}', (prev_out if j == 0 else out_c) + skip_c, out_c)
            if cross_up[i]:
                transformer(f'up_blocks.{i}.attentions.{j}', out_c)
        if i != len(rev) - 1:
            shp[f'up_blocks.{i}.upsamplers.0.conv.weight'] = (out_c, out_c, 3, 3)
            shp[f'up_blocks.{i}.upsamplers.0.conv.bias'] = (out_c,)
    shp['conv_norm_out.weight'] = shp['conv_norm_out.bias'] = (boc[0],)
    shp['conv_out.weight'], shp['conv_out.bias'] = (out_channels, boc[0], 3, 3), (out_channels,)
    return shp


def vae_decoder_param_shapes(latent_channels=4, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                             layers_per_block=2) -> Dict[str, Tuple[int, ...]]:
    """Parameter inventory of the decode half of the SD `AutoencoderKL` (`post_quant_conv` + `decoder.*`): names and shapes
    as the vendored diffusers registers them (sd/diffusers/src/diffusers/models/autoencoders/autoencoder_kl.py:100-130,
    vae.py Decoder :204-290, unet_2d_blocks.py UNetMidBlock2D / UpDecoderBlock2D, resnet.py ResnetBlock2D with temb=None,
    attention_processor.py Attention with one head).  tests/golden/sd_vae_shapes.json pins this against the reference."""
    boc = list(block_out_channels)
    shp: Dict[str, Tuple[int, ...]] = {'post_quant_conv.weight': (latent_channels, latent_channels, 1, 1),
                                       'post_quant_conv.bias': (latent_channels,)}
    top = boc[-1]
    shp['decoder.conv_in.weight'], shp['decoder.conv_in.bias'] = (top, latent_channels, 3, 3), (top,)

    def resnet(p, cin, cout):
        shp[f'{p}.norm1.weight'], shp[f'{p}.norm1.bias'] = (cin,), (cin,)
        shp[f'{p}.conv1.weight'], shp[f'{p}.conv1.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{p}.norm2.weight'], shp[f'{p}.norm2.bias'] = (cout,), (cout,)
        shp[f'{p}.conv2.weight'], shp[f'{p}.conv2.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shp[f'{p}.conv_shortcut.weight'], shp[f'{p}.conv_shortcut.bias'] = (cout, cin, 1, 1), (cout,)

    a = 'decoder.mid_block.attentions.0'
    shp[f'{a}.group_norm.weight'], shp[f'{a}.group_norm.bias'] = (top,), (top,)
    for nm in ('to_q', 'to_k', 'to_v', 'to_out.0'):
        shp[f'{a}.{nm}.weight'], shp[f'{a}.{nm}.bias'] = (top, top), (top,)
    resnet('decoder.mid_block.resnets.0', top, top)
    resnet('decoder.mid_block.resnets.1', top, top)
    rev = boc[::-1]
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(layers_per_block + 1):
            resnet(f'decoder.up_blocks.{i}.resnets.{j}', cin if j == 0 else cout, cout)
        if i != len(rev) - 1:
            shp[f'decoder.up_blocks.{i}.upsamplers.0.conv.weight'] = (cout, cout, 3, 3)
            shp[f'decoder.up_blocks.{i}.upsamplers.0.conv.bias'] = (cout,)
        cin = cout
    shp['decoder.conv_norm_out.weight'], shp['decoder.conv_norm_out.bias'] = (boc[0],), (boc[0],)
    shp['decoder.conv_out.weight'], shp['decoder.conv_out.bias'] = (out_channels, boc[0], 3, 3), (out_channels,)
    return shp


def ddpmpp_param_shapes(img_resolution=32, in_channels=3, out_channels=3, label_dim=0, model_channels=128,
                        channel_mult: Sequence[int] = (2, 2, 2), channel_mult_emb=4, num_blocks=4,
                        attn_resolutions: Sequence[int] = (16,)) -> Dict[str, Tuple[int, ...]]:
    """Parameter inventory of the DDPM++ preset of `SongUNet` (BASELINE.json configs[0]: EDM CIFAR-10 32x32; positional
    embedding, 'standard' encoder / decoder, resample_filter [1,1]): names and shapes as the reference registers them
    (edm/training/networks.py:229-319; preset edm/train.py:118-122).  Differences from ADM: additive embedding (`affine` has
    cout outputs), 1x1 `skip` convs on every resampling block, decoder attention only in the last block of a level, the
    `aux_norm` / `aux_conv` output head.  tests/test_host_logic.py checks it against the oracle's spec-derived inventory."""
    E = model_channels * channel_mult_emb
    shp: Dict[str, Tuple[int, ...]] = {
        'map_layer0.weight': (E, model_channels), 'map_layer0.bias': (E,),
        'map_layer1.weight': (E, E), 'map_layer1.bias': (E,),
    }
    if label_dim:
        shp['map_label.weight'], shp['map_label.bias'] = (model_channels, label_dim), (model_channels,)

    def unet_block(name, cin, cout, attention, resample=False):
        shp[f'{name}.norm0.weight'] = shp[f'{name}.norm0.bias'] = (cin,)
        shp[f'{name}.conv0.weight'], shp[f'{name}.conv0.bias'] = (cout, cin, 3, 3), (cout,)
        shp[f'{name}.affine.weight'], shp[f'{name}.affine.bias'] = (cout, E), (cout,)
        shp[f'{name}.norm1.weight'] = shp[f'{name}.norm1.bias'] = (cout,)
        shp[f'{name}.conv1.weight'], shp[f'{name}.conv1.bias'] = (cout, cout, 3, 3), (cout,)
        if cin != cout or resample:
            shp[f'{name}.skip.weight'], shp[f'{name}.skip.bias'] = (cout, cin, 1, 1), (cout,)
        if attention:
            shp[f'{name}.norm2.weight'] = shp[f'{name}.norm2.bias'] = (cout,)
            shp[f'{name}.qkv.weight'], shp[f'{name}.qkv.bias'] = (3 * cout, cout, 1, 1), (3 * cout,)
            shp[f'{name}.proj.weight'], shp[f'{name}.proj.bias'] = (cout, cout, 1, 1), (cout,)

    skips, c = [], in_channels
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            shp[f'enc.{res}x{res}_conv.weight'], shp[f'enc.{res}x{res}_conv.bias'] = (model_channels, c, 3, 3), (model_channels,)
            c = model_channels
        else:
            unet_block(f'enc.{res}x{res}_down', c, c, False, resample=True)
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
            unet_block(f'dec.{res}x{res}_up', c, c, False, resample=True)
        for idx in range(num_blocks + 1):
            c_new = model_channels * mult
            unet_block(f'dec.{res}x{res}_block{idx}', c + skips.pop(), c_new, idx == num_blocks and res in attn_resolutions)
            c = c_new
        if level == 0:
            shp[f'dec.{res}x{res}_aux_norm.weight'] = shp[f'dec.{res}x{res}_aux_norm.bias'] = (c,)
            shp[f'dec.{res}x{res}_aux_conv.weight'] = (out_channels, c, 3, 3)
            shp[f'dec.{res}x{res}_aux_conv.bias'] = (out_channels,)
    return shp


def clip_param_shapes(hidden=1024, layers=24, intermediate=4096, image_size=224, patch=14, proj=768, t_hidden=768, t_layers=12,
                      t_intermediate=3072, vocab=49408, max_pos=77, vision_only=False) -> Dict[str, Tuple[int, ...]]:
    """Parameter inventory of transformers' `CLIPModel` (the reference's CLIPScorer loads openai/clip-vit-large-patch14,
    sd/scorers.py:150-163; the defaults are that checkpoint's sizes).  `vision_only`: the half the per-candidate path uses.
    tests/test_clip_oracle.py pins the names against the oracle's table, which oracle/make_golden_clip.py checks against
    transformers itself."""
    shp: Dict[str, Tuple[int, ...]] = {}

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
    shp['vision_model.embeddings.position_embedding.weight'] = ((image_size // patch) ** 2 + 1, hidden)
    shp['vision_model.pre_layrnorm.weight'] = shp['vision_model.pre_layrnorm.bias'] = (hidden,)      # sic: transformers' spelling
    tower('vision_model', hidden, intermediate, layers)
    shp['vision_model.post_layernorm.weight'] = shp['vision_model.post_layernorm.bias'] = (hidden,)
    shp['visual_projection.weight'] = (proj, hidden)
    if not vision_only:
        shp['logit_scale'] = ()
        shp['text_model.embeddings.token_embedding.weight'] = (vocab, t_hidden)
        shp['text_model.embeddings.position_embedding.weight'] = (max_pos, t_hidden)
        tower('text_model', t_hidden, t_intermediate, t_layers)
        shp['text_model.final_layer_norm.weight'] = shp['text_model.final_layer_norm.bias'] = (t_hidden,)
        shp['text_projection.weight'] = (proj, t_hidden)
    return shp

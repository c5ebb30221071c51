"""Generates tests/golden/sd_vae_*.{pt,json} by running the VENDORED diffusers AutoencoderKL of the reference
(/root/reference/sd/diffusers) on CPU in THIS container:

  PYTHONHASHSEED=0 python oracle/make_golden_vae.py [--full]

  sd_vae_shapes.json : state-dict names/shapes of the decode half (post_quant_conv + decoder.*), tiny and SD-1.5 configs
  sd_vae_tiny.pt     : AutoencoderKL.decode (tiny config: 64/128 channels, 16x16 latents -> 32x32 image) + per-module taps
  sd_vae_full.pt     : the SD-1.5 decoder (49.5 M parameters), 32x32 latents -> 256x256 image (--full; the 64x64 -> 512x512
                       decode is the same network on 4x the pixels; the fixture stays small)
Weights are never stored: both sides regenerate them with arch.random_state_dict(shapes, seed)."""
import json, os, sys
sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLD = os.path.join(ROOT, 'tests', 'golden')
from make_golden_sd import load_diffusers  # noqa: E402

TINY = dict(block_out_channels=(64, 128), layers_per_block=1)
TINY_KW = dict(in_channels=3, out_channels=3, down_block_types=('DownEncoderBlock2D',) * 2,
               up_block_types=('UpDecoderBlock2D',) * 2, block_out_channels=(64, 128), layers_per_block=1, latent_channels=4,
               norm_num_groups=32, sample_size=32)
FULL_KW = dict(in_channels=3, out_channels=3, down_block_types=('DownEncoderBlock2D',) * 4,
               up_block_types=('UpDecoderBlock2D',) * 4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               latent_channels=4, norm_num_groups=32, sample_size=512)


def decode_half(sd):
    return {k: v for k, v in sd.items() if k.startswith('decoder.') or k.startswith('post_quant_conv.')}


def fixture(AutoencoderKL, kw, shapes, seed, B, h, path, keep_names=None):
    from diffusion_tts_b200.arch import random_state_dict
    vae = AutoencoderKL(**kw).eval().requires_grad_(False)
    sd = random_state_dict(shapes, seed)
    missing, unexpected = vae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(('encoder.', 'quant_conv.')) for k in missing), (missing[:4], unexpected[:4])
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn(B, 4, h, h, generator=g)
    taps, hooks = {}, []
    for name, mod in vae.named_modules():
        leaf = name.split('.')
        if name.startswith('decoder.') and ((len(leaf) == 5 and leaf[3] == 'resnets') or (len(leaf) == 4 and leaf[1] == 'mid_block')
                                            or name == 'decoder.conv_in' or name.endswith('upsamplers.0.conv')):
            hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        img = vae.decode(z, return_dict=False)[0]
    for hk in hooks:
        hk.remove()
    keep = {k: v.to(torch.float16) for k, v in taps.items() if keep_names is None or k in keep_names}
    torch.save(dict(seed=seed, z=z, out=img, taps=keep, tap_absmean={k: float(v.abs().mean()) for k, v in taps.items()}), path)
    print(path, tuple(img.shape), float(img.abs().mean()), len(keep), 'of', len(taps), 'taps kept')


def main():
    load_diffusers()
    from diffusers import AutoencoderKL
    from diffusion_tts_b200.arch import vae_decoder_param_shapes
    with torch.device('meta'):
        ref_full = {k: list(v.shape) for k, v in decode_half(AutoencoderKL(**FULL_KW).state_dict()).items()}
        ref_tiny = {k: list(v.shape) for k, v in decode_half(AutoencoderKL(**TINY_KW).state_dict()).items()}
    json.dump(dict(full=ref_full, tiny=ref_tiny), open(os.path.join(GOLD, 'sd_vae_shapes.json'), 'w'))
    assert {k: tuple(v) for k, v in ref_full.items()} == vae_decoder_param_shapes()
    assert {k: tuple(v) for k, v in ref_tiny.items()} == vae_decoder_param_shapes(**TINY)
    fixture(AutoencoderKL, TINY_KW, vae_decoder_param_shapes(**TINY), 41, 2, 16, os.path.join(GOLD, 'sd_vae_tiny.pt'))
    if '--full' in sys.argv:
        fixture(AutoencoderKL, FULL_KW, vae_decoder_param_shapes(), 42, 1, 32, os.path.join(GOLD, 'sd_vae_full.pt'),
                keep_names=('decoder.conv_in', 'decoder.mid_block.attentions.0', 'decoder.mid_block.resnets.1'))


if __name__ == '__main__':
    main()

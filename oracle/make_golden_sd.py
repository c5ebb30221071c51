"""Generates tests/golden/sd_*.{pt,json} by running the VENDORED diffusers modules of the reference
(/root/reference/sd/diffusers, loaded the way the reference's main.py:48-51 does) on CPU in THIS container:

  PYTHONHASHSEED=0 python oracle/make_golden_sd.py [--full]

  sd_unet_shapes.json : state-dict names/shapes of the tiny and the SD-1.5-shaped UNet2DConditionModel
  sd_unet_tiny.pt     : UNet2DConditionModel forward (tiny config) on seeded weights / inputs + per-module taps
  sd_unet_full.pt     : the same for the SD-1.5-shaped model (859.5 M parameters; --full, ~1 min, ~8 GB)
  sd_ddim.pt          : DDIMScheduler (reference-edited step: eta=1 default, tuple return) timesteps and step outputs
Weights are never stored: both sides regenerate them with arch.random_state_dict(shapes, seed)."""
import importlib.util, json, os, sys
sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def load_diffusers():
    import transformers.utils
    if not hasattr(transformers.utils, 'FLAX_WEIGHTS_NAME'):
        transformers.utils.FLAX_WEIGHTS_NAME = 'flax_model.msgpack'
    root = '/root/reference/sd/diffusers/src/diffusers'
    spec = importlib.util.spec_from_file_location('diffusers', root + '/__init__.py', submodule_search_locations=[root])
    m = importlib.util.module_from_spec(spec)
    sys.modules['diffusers'] = m
    spec.loader.exec_module(m)
    import types
    from diffusers import DDIMScheduler, UNet2DConditionModel           # the package swaps itself for a lazy module
    return types.SimpleNamespace(UNet2DConditionModel=UNet2DConditionModel, DDIMScheduler=DDIMScheduler)


TINY = dict(block_out_channels=(64, 128), layers_per_block=1, cross_attn_down=(True, False), cross_attention_dim=64)
TINY_KW = dict(sample_size=16, block_out_channels=(64, 128), layers_per_block=1,
               down_block_types=('CrossAttnDownBlock2D', 'DownBlock2D'), up_block_types=('UpBlock2D', 'CrossAttnUpBlock2D'),
               cross_attention_dim=64)
FULL_KW = dict(sample_size=64, cross_attention_dim=768)


def unet_fixture(D, kw, shapes, seed, B, H, ctx_dim, t, path, max_tap=2 ** 21):
    from diffusion_tts_b200.arch import random_state_dict
    net = D.UNet2DConditionModel(**kw).eval().requires_grad_(False)
    sd = random_state_dict(shapes, seed)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 4, H, H, generator=g)
    ctx = torch.randn(B, 77, ctx_dim, generator=g)
    taps = {}
    hooks = []
    for name, mod in net.named_modules():
        leaf = name.split('.')
        if (len(leaf) == 4 and leaf[2] in ('resnets', 'attentions')) or (len(leaf) == 3 and leaf[0] == 'mid_block') or \
                name == 'conv_in' or name.endswith(('downsamplers.0.conv', 'upsamplers.0.conv')):
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, name=name: taps.__setitem__(name, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    with torch.no_grad():
        out = net(x, t, encoder_hidden_states=ctx, return_dict=False)[0]
    for h in hooks:
        h.remove()
    keep = {k: v for k, v in taps.items() if v.numel() <= max_tap}          # keep the fixture small
    torch.save(dict(seed=seed, x=x, ctx=ctx, t=t, out=out, taps={k: v.to(torch.float16) for k, v in keep.items()},
                    tap_absmean={k: float(v.abs().mean()) for k, v in taps.items()}), path)
    print(path, tuple(out.shape), float(out.abs().mean()), len(keep), 'taps')


def main():
    D = load_diffusers()
    from diffusion_tts_b200.arch import sd_unet_param_shapes
    with torch.device('meta'):
        ref_full = {k: list(v.shape) for k, v in D.UNet2DConditionModel(**FULL_KW).state_dict().items()}
        ref_tiny = {k: list(v.shape) for k, v in D.UNet2DConditionModel(**TINY_KW).state_dict().items()}
    json.dump(dict(full=ref_full, tiny=ref_tiny), open(os.path.join(GOLD, 'sd_unet_shapes.json'), 'w'))
    assert {k: tuple(v) for k, v in ref_full.items()} == sd_unet_param_shapes()
    assert {k: tuple(v) for k, v in ref_tiny.items()} == sd_unet_param_shapes(**TINY)
    unet_fixture(D, TINY_KW, sd_unet_param_shapes(**TINY), 11, 4, 16, 64, 500, os.path.join(GOLD, 'sd_unet_tiny.pt'))
    if '--full' in sys.argv:
        unet_fixture(D, FULL_KW, sd_unet_param_shapes(), 12, 2, 64, 768, 621, os.path.join(GOLD, 'sd_unet_full.pt'),
                     max_tap=2 ** 18)
    # ---- scheduler: SD-1.5 scheduler config
    sch = D.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear', clip_sample=False,
                          set_alpha_to_one=False, steps_offset=1)
    sch.set_timesteps(10)
    g = torch.Generator().manual_seed(5)
    rows = []
    for t in [int(v) for v in sch.timesteps]:
        eps, sample, noise = [torch.randn(2, 4, 8, 8, generator=g) for _ in range(3)]
        prev, x0 = sch.step(eps, t, sample, variance_noise=noise, return_dict=False)
        rows.append(dict(t=t, eps=eps, sample=sample, noise=noise, prev=prev, x0=x0))
    torch.save(dict(num_inference_steps=10, timesteps=[int(v) for v in sch.timesteps], rows=rows,
                    alphas_cumprod=sch.alphas_cumprod.clone()), os.path.join(GOLD, 'sd_ddim.pt'))
    print('sd_ddim.pt', [int(v) for v in sch.timesteps])


if __name__ == '__main__':
    main()

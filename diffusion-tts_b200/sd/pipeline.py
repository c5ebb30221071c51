"""Host-side mirror of the reference SD entry point for the beam step (BASELINE.json config 5).

The reference drives the SD backend as
    result, score = pipe(prompt=..., num_inference_steps=50, score_function=scorer, method=method, params=MASTER_PARAMS)
(main.py:135-141; StableDiffusionPipeline.__call__, pipeline_stable_diffusion.py:785-816, returns `(out, score)` :1485).
`B200LatentBeamPipeline.__call__` keeps that call shape for `method="beam"` and runs the search on the B200 engine
(`sd_beam_search`).  What is deliberately different, all forced by config 5 / the offline setting:
  * without a VAE, scoring happens in latent space on the Tweedie x0 (config 5: "brightness scorer on Tweedie x0");
    with `vae_state_dict` every candidate's x0 is decoded to an image on the B200 VAE engine (vae.py) and
    the image is scored, as the reference does (pipeline_stable_diffusion.py:1111-1123);
  * the CLIP text encoder's weights are unreachable offline: `encode_prompt` is injectable, and the default produces a
    deterministic pseudo-embedding pair [uncond, cond] of the right shape from the prompt string;
  * `out.images` is the decoded image when a VAE is given, else a visualisation of the first three latent channels;
    `out.latents` is the search result itself.
`method` = beam (sd/beam.py), eps_greedy / zero_order / naive (sd/search.py; 'rejection' is the naive loop, repeated by
main.py like the reference's main.py:131), mcts (sd/search.py:sd_mcts_as_shipped -- the reference branch as shipped never
back-propagates a reward, so its observable behaviour is a DDIM step with the step's first noise draw).
"""
from __future__ import annotations

import hashlib
from types import SimpleNamespace
from typing import Callable, Dict, Optional

import torch

from ..sd_unet import SDUNetEngine
from .beam import DDIMTable, sd_beam_search
from .search import sd_eps_greedy_search, sd_mcts_as_shipped


def pseudo_prompt_embeddings(prompt: str, negative_prompt: str = '', tokens: int = 77, dim: int = 768) -> torch.Tensor:
    """[2, tokens, dim] = [negative/uncond, prompt] stand-ins for CLIP last_hidden_state (layer-normed scale ~ 1)."""
    out = []
    for text in (negative_prompt, prompt):
        seed = int.from_bytes(hashlib.sha256(text.encode()).digest()[:8], 'little') % (2 ** 63)
        out.append(torch.randn(tokens, dim, generator=torch.Generator().manual_seed(seed)))
    return torch.stack(out)


class B200LatentBeamPipeline:
    def __init__(self, unet_state_dict: Dict[str, torch.Tensor], device='cuda', encode_prompt: Optional[Callable] = None,
                 decode: Optional[Callable] = None, shard=None, vae_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 vae_chunk: int = 16):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('B200LatentBeamPipeline runs on a B200 only (no CPU fallback)')
        self.unet = SDUNetEngine(unet_state_dict, device=self.device)
        self.encode_prompt = encode_prompt or (lambda p, n='': pseudo_prompt_embeddings(p, n, dim=self.unet.cfg['cross_attention_dim']))
        self.decode = decode
        self.shard = shard
        self.vae = None
        if vae_state_dict is not None:
            from ..vae import VAEDecoderEngine
            self.vae = VAEDecoderEngine(vae_state_dict, device=self.device)
        self.vae_chunk = vae_chunk

    @torch.no_grad()
    def __call__(self, prompt: str, num_inference_steps: int = 50, score_function: Optional[Callable] = None,
                 method: str = 'beam', params: Optional[dict] = None, guidance_scale: float = 7.5,
                 negative_prompt: str = '', latents: Optional[torch.Tensor] = None, height: int = 512, width: int = 512,
                 generator: Optional[torch.Generator] = None):
        if method not in ('beam', 'eps_greedy', 'zero_order', 'naive', 'rejection', 'mcts'):
            raise ValueError(f"Unknown method: {method}")
        params = params or {}
        if height != width or height % 64:
            raise ValueError('height == width, a multiple of 64, is required')
        H = height // 8                                                   # vae_scale_factor = 8
        if latents is None:                                               # prepare_latents: randn * init_noise_sigma (= 1)
            latents = torch.randn(1, self.unet.cfg['in_channels'], H, H, generator=generator)
        ctx = self.encode_prompt(prompt, negative_prompt)
        table = DDIMTable(num_inference_steps)
        fused = self.decode is None and (score_function is None or getattr(score_function, 'latent_fused', False))
        kw = dict(guidance_scale=guidance_scale, shard=self.shard)
        if self.vae is not None and self.decode is None:      # decode + quantise + score every candidate's x0 (:1111-1123)
            from ..vae import DecodedImageScorer
            kw['decode'] = lambda x0: x0
            kw['scorer'] = DecodedImageScorer(self.vae, score_function, prompt, self.vae_chunk)
        elif not fused:
            kw['decode'] = self.decode or (lambda x0: (x0 * 127.5 + 128).clip(0, 255).to(torch.uint8))    # :1115
            kw['scorer'] = lambda im: score_function(im, [prompt] * im.shape[0], torch.zeros(im.shape[0], device=im.device))
        if method == 'beam':
            B, N = int(params['B']), int(params['N'])                      # pipeline_stable_diffusion.py:1046,1080
            best, rec = sd_beam_search(self.unet, table, latents, ctx, B, N, **kw)
            score = float(rec.final_score)
        elif method == 'mcts':   # as shipped: no reward ever reaches the tree (see sd_mcts_as_shipped); same RNG consumption
            kw.pop('shard')
            best, rec = sd_mcts_as_shipped(self.unet, table, latents, ctx, int(params['N']), int(params['S']), **kw)
            score = float(rec.max_score)
        else:   # eps_greedy / zero_order, and the plain eta=1 DDIM loop every other method name falls through to (:1330-1437)
            m = method if method in ('eps_greedy', 'zero_order') else 'naive'
            best, rec = sd_eps_greedy_search(self.unet, table, latents, ctx, int(params.get('N', 1)), int(params.get('K', 1)),
                                             float(params.get('lambda', 0.15)), float(params.get('eps', 0.4)), m, **kw)
            score = float(rec.max_score)
        from PIL import Image
        if self.vae is not None:
            img = self.vae.decode_latents(best)[0]                                                  # fp32 [H, W, 3]
            vis = (img * 127.5 + 128).clip(0, 255).to(torch.uint8).cpu().numpy()
        else:
            vis = (best[0, :3] * 127.5 + 128).clip(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        out = SimpleNamespace(images=[Image.fromarray(vis, 'RGB')], latents=best, record=rec)
        return out, score

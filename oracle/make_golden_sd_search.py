"""Pins the SD search LOOPS of the oracle (oracle/sd_oracle.py: beam_search, eps_greedy_search and the reference's RNG
call sequence) against the REAL edited pipeline: assembles a tiny random-init `StableDiffusionPipeline` from the vendored
diffusers of the reference (/root/reference/sd/diffusers; UNet2DConditionModel + AutoencoderKL + DDIMScheduler, no text
encoder: prompt embeddings are passed in) and runs its `__call__` (pipeline_stable_diffusion.py:785-1485) with
`method = 'beam' | 'eps_greedy' | 'zero_order'` and a recording `score_function` on CPU:

  PYTHONHASHSEED=0 python oracle/make_golden_sd_search.py        ->  tests/golden/sd_search_tiny.pt

Stored: seeds, the score of every scored candidate in call order, the returned latents and max_score.  Weights and inputs
are regenerated from the seeds by the test (tests/test_sd_oracle.py); noise is NOT stored -- the oracle must re-draw it
from the seed in the reference's own call order (incl. the scheduler's discarded variance-noise draw of the scoring-only
second step, scheduling_ddim.py:457-461), which pins that order too.  TEST INFRASTRUCTURE ONLY."""
import os
import sys

sys.dont_write_bytecode = True
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLD = os.path.join(ROOT, 'tests', 'golden')
from make_golden_sd import TINY as UNET_TINY, TINY_KW as UNET_KW, load_diffusers  # noqa: E402
from make_golden_vae import TINY as VAE_TINY, TINY_KW as VAE_KW  # noqa: E402

STEPS, H = 4, 16
UNET_SEED, VAE_SEED, INPUT_SEED = 11, 41, 77


def build_pipe():
    load_diffusers()
    from diffusers import AutoencoderKL, DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel
    from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes, vae_decoder_param_shapes
    unet = UNet2DConditionModel(**UNET_KW).eval().requires_grad_(False)
    unet.load_state_dict(random_state_dict(sd_unet_param_shapes(**UNET_TINY), UNET_SEED), strict=True)
    vae = AutoencoderKL(**VAE_KW).eval().requires_grad_(False)
    vae.load_state_dict(random_state_dict(vae_decoder_param_shapes(**VAE_TINY), VAE_SEED), strict=False)
    sch = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear', clip_sample=False,
                        set_alpha_to_one=False, steps_offset=1)
    pipe = StableDiffusionPipeline(vae=vae, text_encoder=None, tokenizer=None, unet=unet, scheduler=sch, safety_checker=None,
                                   feature_extractor=None, requires_safety_checker=False)
    pipe.set_progress_bar_config(disable=True)
    return pipe


def inputs():
    g = torch.Generator().manual_seed(INPUT_SEED)
    latents = torch.randn(1, 4, H, H, generator=g)
    ctx = torch.randn(2, 77, UNET_KW['cross_attention_dim'], generator=g)         # [uncond, cond]
    return latents, ctx


class Rec:
    def __init__(self):
        self.scores = []

    def __call__(self, images, prompts, timesteps):
        import importlib.util
        img = images[0]                                                         # uint8 [1,3,h,w] (pipeline...:1115)
        w = torch.tensor([0.2126, 0.7152, 0.0722]).view(1, 3, 1, 1)
        if img.shape[1] == 3:
            s = ((img.float() / 255.0) * w).sum(dim=1).mean(dim=(1, 2)).clamp(0, 1)   # sd/scorers.py:43-64, RGB branch
        else:                           # output_type='latent': the final 4-channel latents are "the image" (:1463, :66-67)
            s = (img.float() / 255.0).mean(dim=(1, 2, 3))
        self.scores.append(float(s))
        return s


def run(pipe, method, params, seed):
    latents, ctx = inputs()
    rec = Rec()
    torch.manual_seed(seed)
    out, max_score = pipe(prompt=None, prompt_embeds=ctx[1:2], negative_prompt_embeds=ctx[0:1], latents=latents.clone(),
                          num_inference_steps=STEPS, guidance_scale=7.5, score_function=rec, method=method, params=params,
                          output_type='latent')
    return dict(method=method, params=params, seed=seed, scores=torch.tensor(rec.scores), latents=out.images.clone(),
                max_score=float(max_score))


def main():
    assert os.environ.get('PYTHONHASHSEED') == '0', 'run with PYTHONHASHSEED=0'
    pipe = build_pipe()
    cases = [run(pipe, 'beam', {'B': 2, 'N': 3, 'K': 1, 'lambda': 0.15, 'eps': 0.4, 'S': 8}, 5),
             run(pipe, 'eps_greedy', {'B': 2, 'N': 4, 'K': 2, 'lambda': 0.15, 'eps': 0.4, 'S': 8}, 6),
             run(pipe, 'zero_order', {'B': 2, 'N': 3, 'K': 1, 'lambda': 0.15, 'eps': 0.4, 'S': 8}, 7),
             run(pipe, 'mcts', {'B': 2, 'N': 2, 'K': 1, 'lambda': 0.15, 'eps': 0.4, 'S': 3}, 8)]
    path = os.path.join(GOLD, 'sd_search_tiny.pt')
    torch.save(dict(steps=STEPS, H=H, unet_seed=UNET_SEED, vae_seed=VAE_SEED, input_seed=INPUT_SEED, guidance=7.5,
                    scaling_factor=float(pipe.vae.config.scaling_factor), cases=cases), path)
    for c in cases:
        print(c['method'], 'scored', len(c['scores']), 'max_score', c['max_score'])
    print('wrote', path, os.path.getsize(path), 'B')


if __name__ == '__main__':
    main()

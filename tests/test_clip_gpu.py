"""CLIP scorer (reference sd/scorers.py:149-213) on a B200 against transformers' own outputs
(tests/golden/clip_*.pt from oracle/make_golden_clip.py) and the oracle.

Bars: the preprocessing (Pillow bicubic + crop + rescale + normalise) is integer/table work -> BIT-EXACT; the ViT tower
computes in 16-bit storage with fp32 accumulation -> image embedding within 2e-2 relative (L2) and the cosine score within
5e-3 absolute of the fp32 reference (measured values are printed)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clip_oracle as CO  # noqa: E402
from tests.helpers import load_golden  # noqa: E402

KEYS = ('hidden', 'layers', 'intermediate', 'image_size', 'patch', 'proj', 't_hidden', 't_layers', 't_intermediate', 'vocab', 'max_pos')
EMB_REL, SCORE_ABS = 2e-2, 5e-3


@pytest.fixture(scope='module')
def clip():
    from diffusion_tts_b200 import build
    build.build()
    import diffusion_tts_b200.clip as c
    return c


def _sd(fx):
    return CO.seeded_clip_state_dict(CO.clip_param_shapes(**{k: fx['cfg'][k] for k in KEYS}), fx['seed'])


def _patch_matrix(pv, P, Lp, Kp, dtype):
    """pixel_values fp32 [B,3,S,S] -> the A matrix the preprocessing kernel must produce."""
    B, _, S, _ = pv.shape
    G = S // P
    x = pv.view(B, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(B, G * G, 3 * P * P)
    out = torch.zeros(B, Lp, Kp)
    out[:, 1:1 + G * G, :3 * P * P] = x
    return out.view(B * Lp, Kp).to(dtype)


def test_preprocess_is_bit_exact_and_tiny_tower_matches(clip):
    from diffusion_tts_b200._lib import ACT_DTYPE
    fx = load_golden('clip_tiny.pt')
    sd = _sd(fx)
    scorer = clip.CLIPScorer(sd, device='cuda', text_heads=fx['t_heads'])
    scorer.set_text_embeds('p', fx['text_embeds'])
    img = fx['images'].cuda()
    s = scorer(img, ['p'], None).cpu()
    eng = scorer.engine
    cp = eng.plan(*[img.shape[i] for i in (0, 2, 3)])
    want = _patch_matrix(fx['pixel_values'], eng.cfg['patch'], eng.Lp, eng.Kp, ACT_DTYPE)
    assert torch.equal(cp.patches.cpu(), want)                                            # bit-exact preprocessing
    rel = float((cp.embeds.cpu() - fx['image_embeds']).norm() / fx['image_embeds'].norm())
    print('tiny: embeds rel', rel, 'score', s.tolist(), 'ref', fx['score'].tolist())
    assert rel < EMB_REL
    assert torch.allclose(s, fx['score'], rtol=0, atol=SCORE_ABS)
    # the text tower (once per prompt, torch fp32 on the GPU) against transformers
    te = clip.encode_text(sd, fx['input_ids'], fx['t_heads'], 'cuda').cpu()
    assert torch.allclose(te, fx['text_embeds'], rtol=1e-3, atol=1e-4)
    scorer2 = clip.CLIPScorer(sd, device='cuda', text_heads=fx['t_heads'], tokenize=lambda p: fx['input_ids'])
    assert torch.allclose(scorer2(img, 'anything', None).cpu(), s, rtol=0, atol=1e-4)
    # protocol edges of the reference scorer: prompts None -> zeros; identical images -> identical bits; batch-position invariance
    assert torch.equal(scorer(img, None, None).cpu(), torch.zeros(img.shape[0]))
    two = torch.stack([img[1], img[1], img[0]])
    s2 = scorer(two, ['p'] * 3, None).cpu()
    assert s2[0] == s2[1] and s2[2] == s[0] and s2[0] == s[1]
    with pytest.raises(RuntimeError):
        scorer(img, ['unknown prompt'], None)
    with pytest.raises(TypeError):
        scorer(img.float(), ['p'], None)


def test_ragged_input_sizes_are_bit_exact(clip):
    """Non-square, up- and down-scaled inputs: resize of the shortest edge + centre crop, against the oracle (itself pinned
    against Pillow on the CPU)."""
    from diffusion_tts_b200._lib import ACT_DTYPE
    fx = load_golden('clip_tiny.pt')
    scorer = clip.CLIPScorer(_sd(fx), device='cuda', text_heads=fx['t_heads'])
    scorer.set_text_embeds('p', fx['text_embeds'])
    eng = scorer.engine
    g = torch.Generator().manual_seed(5)
    for (h, w) in ((128, 96), (56, 56), (40, 75), (200, 333)):
        img = torch.randint(0, 256, (2, 3, h, w), generator=g, dtype=torch.uint8)
        scorer(img.cuda(), ['p'], None)
        want = _patch_matrix(CO.clip_preprocess(img, eng.cfg['image_size']), eng.cfg['patch'], eng.Lp, eng.Kp, ACT_DTYPE)
        assert torch.equal(eng.plan(2, h, w).patches.cpu(), want), (h, w)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), 'golden', 'clip_vitl14.pt')), reason='full fixture not generated')
def test_vitl14_matches_transformers(clip):
    """openai/clip-vit-large-patch14's architecture (seeded weights), 512 x 512 inputs as the SD pipeline produces."""
    from oracle.make_golden_clip import images
    fx = load_golden('clip_vitl14.pt')
    scorer = clip.CLIPScorer(_sd(fx), device='cuda', text_heads=fx['t_heads'])
    scorer.set_text_embeds('p', fx['text_embeds'])
    img = images(fx['image_seed'], fx['n'], *fx['image_hw'])
    s = scorer(img.cuda(), ['p'], None).cpu()
    cp = scorer.engine.plan(fx['n'], *fx['image_hw'])
    rel = float((cp.embeds.cpu() - fx['image_embeds']).norm() / fx['image_embeds'].norm())
    print('ViT-L/14: embeds rel', rel, 'score', s.tolist(), 'ref', fx['score'].tolist())
    assert rel < EMB_REL
    assert torch.allclose(s, fx['score'], rtol=0, atol=SCORE_ABS)
    # larger batch, same images: per-image results do not depend on the batch size
    s8 = scorer(torch.cat([img, img, img, img]).cuda(), ['p'], None).cpu()
    assert torch.equal(s8[:2], s8[2:4]) and torch.allclose(s8[:2], s, rtol=0, atol=1e-4)


def test_clip_scorer_in_the_sd_search_loop(clip):
    """SD eps_greedy with decode-then-score and the CLIP scorer as `score_function` (reference main.py:66-67 + 135-141) on
    tiny networks: every uint8 image the loop hands to the scorer is re-scored by the fp32 oracle (Pillow-exact
    preprocessing + ViT); the loop's selection must be the argmax of the scores it was given."""
    import diffusion_tts_b200.arch as arch
    from diffusion_tts_b200.sd.pipeline import B200LatentBeamPipeline
    fx = load_golden('clip_tiny.pt')
    csd = _sd(fx)
    inner = clip.CLIPScorer(csd, device='cuda', text_heads=fx['t_heads'])
    inner.set_text_embeds('a photo', fx['text_embeds'])
    seen = []

    def recording(images, prompts, timesteps):
        s = inner(images, prompts, timesteps)
        seen.append((images.cpu().clone(), list(prompts), s.cpu().clone()))
        return s

    UT = dict(block_out_channels=(64, 128), layers_per_block=1, cross_attn_down=(True, False), cross_attention_dim=64)
    usd = arch.random_state_dict(arch.sd_unet_param_shapes(**UT), 51)
    vsd = arch.random_state_dict(arch.vae_decoder_param_shapes(block_out_channels=(64, 128), layers_per_block=1), 52)
    pipe = B200LatentBeamPipeline(usd, device='cuda', vae_state_dict=vsd)
    torch.manual_seed(3)
    out, score = pipe(prompt='a photo', num_inference_steps=2, score_function=recording, method='eps_greedy',
                      params={'N': 3, 'K': 2, 'lambda': 0.15, 'eps': 0.4}, height=128, width=128)
    assert len(seen) >= 4 and sum(im.shape[0] for im, _, _ in seen) == 2 * 2 * 3
    worst = 0.0
    for im, prompts, s in seen:
        assert im.dtype == torch.uint8 and im.shape[1] == 3 and prompts == ['a photo'] * im.shape[0]
        with torch.no_grad():
            ref = CO.clip_score(CO.clip_vision_forward(csd, CO.clip_preprocess(im, fx['cfg']['image_size']), fx['v_heads']),
                                fx['text_embeds'])
        worst = max(worst, float((ref - s).abs().max()))
    print('CLIP in the SD loop: worst |score - oracle| =', worst)
    assert worst < SCORE_ABS
    assert abs(score - float(seen[-1][2].max())) < 1e-6

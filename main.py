"""Command-line front end with the reference's flags (reference main.py:80-98), driving the
B200 noise-search path behind `--method/--scorer` for the EDM backend.

    python main.py --backend edm --scorer brightness --method eps_greedy --N 64 --K 20

Differences from the reference CLI, all forced by the offline B200 setting:
  * `--network` (new, optional): a local network pickle / `.pt` bundle.  The reference downloads
    the ImageNet-64 ADM pickle (main.py:157-158); without `--network` this front end builds a
    random-init ADM of the same architecture so that the path can run without network access.
  * `--arch adm|ddpmpp` (new, optional): which random-init architecture stands in when `--network` is absent -- ddpmpp is the
    CIFAR-10 32x32 SongUNet of BASELINE.json configs[0] (`--method naive`, unconditional).
  * `--classifier` (new, optional): a local `64x64_classifier.pt` for `--scorer imagenet` (random-init otherwise).
  * `--backend sd` runs `--method beam | eps_greedy | zero_order | naive | rejection` (BASELINE.json config 5 = beam): an SD-1.5-shaped UNet (random-init unless
    `--network` names a state-dict `.pt`), latent-space scoring of the Tweedie x0, pseudo prompt embeddings (the CLIP
    text encoder and the VAE are unreachable offline).  `--steps` (new) overrides the 50 DDIM steps;
    `--vae PATH|random` (new) decodes every candidate's x0 to a 512x512 image on the B200 VAE engine before scoring.
  * `--scorer clip` (sd): the ViT tower + Pillow-exact preprocessing run on the B200 engine (diffusion_tts_b200/clip.py);
    `--clip PATH|random` (new) is a transformers CLIPModel state dict (random-init ViT-L/14 otherwise -- the reference downloads
    openai/clip-vit-large-patch14), `--clip-text PATH` (new) a precomputed prompt embedding [1, 768]; without it (no tokenizer
    vocabulary offline) a deterministic pseudo embedding of the prompt string stands in.  Needs decoded images: implies `--vae`.
  * `--method mcts`: edm runs batched expansions / rollouts; sd runs the reference branch as shipped.
"""
import argparse

import torch


def get_scorer(backend, scorer_name, device='cuda', classifier=None):
    """Scorer factory (reference main.py:60-71)."""
    from diffusion_tts_b200 import scorers
    if scorer_name == 'brightness':
        return scorers.BrightnessScorer(dtype=torch.float32, device=device)
    if scorer_name == 'imagenet' and backend == 'edm':
        from diffusion_tts_b200.arch import classifier_param_shapes, random_state_dict
        from diffusion_tts_b200.classifier import ImageNetScorer
        if classifier is not None:
            sd = torch.load(classifier, map_location='cpu')
        else:       # the reference downloads 64x64_classifier.pt (edm/scorers.py:61-74): unreachable offline
            print('[imagenet scorer] no --classifier checkpoint given: using a random-init ADM classifier')
            sd = random_state_dict(classifier_param_shapes(), 22)
        return ImageNetScorer(sd, dtype=torch.float32, device=device)
    if scorer_name == 'compressibility':
        return scorers.CompressibilityScorer(dtype=torch.float32, device=device)
    raise ValueError(f"Unknown or invalid scorer '{scorer_name}' for backend '{backend}'")


def get_clip_scorer(args):
    """CLIPScorer (reference main.py:66-67, sd/scorers.py:149-213) on the B200 engine."""
    import hashlib
    from diffusion_tts_b200.arch import clip_param_shapes, random_state_dict
    from diffusion_tts_b200.clip import CLIPScorer
    if args.clip not in (None, 'random'):
        sd = torch.load(args.clip, map_location='cpu')
    else:
        print('[clip scorer] no --clip checkpoint given: using a random-init CLIP ViT-L/14 vision tower (303.5M parameters)')
        sd = random_state_dict(clip_param_shapes(vision_only=True), 33)
    scorer = CLIPScorer(sd, dtype=torch.float32, device=args.device)
    if args.clip_text is not None:
        emb = torch.load(args.clip_text, map_location='cpu')
    else:
        print('[clip scorer] no --clip-text embedding given: using a pseudo embedding of the prompt string')
        seed = int.from_bytes(hashlib.sha256(args.prompt.encode()).digest()[:8], 'little') % (2 ** 63)
        emb = torch.randn(1, scorer.engine.cfg['proj'], generator=torch.Generator().manual_seed(seed))
    scorer.set_text_embeds(args.prompt, emb)
    return scorer


def random_init_ddpmpp(seed=4321):
    """CIFAR-10 32x32 DDPM++ (SongUNet) bundle, random init (BASELINE.json configs[0]; reference preset edm/train.py:118-122)."""
    from diffusion_tts_b200.arch import ddpmpp_param_shapes, random_state_dict
    return dict(state_dict=random_state_dict(ddpmpp_param_shapes(), seed), sigma_data=0.5)


def random_init_adm(seed=1234):
    """ImageNet-64 ADM bundle with every tensor drawn N(0, 1/fan_in) (the reference's zero-initialised
    conv1/proj/out_conv would make F_x == 0, SURVEY.md 7 hard part 2)."""
    from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
    return dict(state_dict=random_state_dict(adm_param_shapes(), seed), sigma_data=0.5)


def main_sd(args):
    """SD backend (reference main.py:111-146): `pipe(prompt, num_inference_steps=50, score_function, method, params)`."""
    from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes
    from diffusion_tts_b200.sd.pipeline import B200LatentBeamPipeline
    if args.scorer == 'clip':
        scorer = get_clip_scorer(args)
        args.vae = args.vae or 'random'            # CLIP scores decoded RGB images
    else:
        scorer = get_scorer('sd', args.scorer, args.device)
    if args.network is not None:
        sd = torch.load(args.network, map_location='cpu')
    else:
        print('[sd] no --network given: using a random-init SD-1.5-shaped UNet2DConditionModel (859.5M parameters)')
        sd = random_state_dict(sd_unet_param_shapes(), 1234)
    torch.manual_seed(args.seed)
    vae_sd = None
    if args.vae is not None:        # decode every candidate's Tweedie x0 to an image before scoring, like the reference
        if args.vae == 'random':
            from diffusion_tts_b200.arch import vae_decoder_param_shapes
            print('[sd] --vae random: random-init SD-1.5 AutoencoderKL decoder (49.5M parameters)')
            vae_sd = random_state_dict(vae_decoder_param_shapes(), 4321)
        else:
            vae_sd = torch.load(args.vae, map_location='cpu')
    pipe = B200LatentBeamPipeline(sd, device=args.device, vae_state_dict=vae_sd)
    params = {'N': args.N, 'lambda': args.lambda_, 'eps': args.eps, 'K': args.K, 'B': args.B, 'S': args.S}
    best_result, best_score = None, float('-inf')
    for _ in range(params['N'] if args.method == 'rejection' else 1):
        result, score = pipe(prompt=args.prompt, num_inference_steps=args.steps or 50, score_function=scorer,
                             method=args.method, params=params)
        if score > best_score:
            best_result, best_score = result, score
    outname = args.output or f"sd_{args.method}_{args.scorer}.png"
    best_result.images[0].save(outname)
    print(f"\n[SD] Saved: {outname}\nBest score: {best_score}\n")


def main():
    parser = argparse.ArgumentParser(description='Unified Diffusion Image Generator (EDM/SD) -- B200 path',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--backend', type=str, choices=['edm', 'sd'], required=True, help='Backend: edm or sd')
    parser.add_argument('--scorer', type=str, choices=['brightness', 'compressibility', 'clip', 'imagenet'],
                        required=True, help='Scorer name')
    parser.add_argument('--method', type=str, default='naive',
                        help='Sampling method (naive, rejection, beam, mcts, zero_order, eps_greedy)')
    parser.add_argument('--prompt', type=str, default='YOUR PROMPT HERE', help='Prompt for SD')
    parser.add_argument('--output', type=str, default=None, help='Output filename (default: auto)')
    parser.add_argument('--N', type=int, default=4, help='Master param N')
    parser.add_argument('--lambda_', type=float, default=0.15, help='Master param lambda')
    parser.add_argument('--eps', type=float, default=0.4, help='Master param eps')
    parser.add_argument('--K', type=int, default=20, help='Master param K')
    parser.add_argument('--B', type=int, default=2, help='Master param B')
    parser.add_argument('--S', type=int, default=8, help='Master param S')
    parser.add_argument('--seed', type=int, default=0, help='Random seed')
    parser.add_argument('--device', type=str, default='cuda', help='Device')
    parser.add_argument('--network', type=str, default=None, help='Local network pickle / .pt bundle')
    parser.add_argument('--classifier', type=str, default=None, help='Local 64x64_classifier.pt for --scorer imagenet')
    parser.add_argument('--arch', type=str, default='adm', choices=['adm', 'ddpmpp'],
                        help='edm, without --network: random-init ImageNet-64 ADM (class-cond) or CIFAR-10 32x32 DDPM++ (uncond)')
    parser.add_argument('--steps', type=int, default=None, help='Override the number of sampler steps (sd: 50)')
    parser.add_argument('--vae', type=str, default=None,
                        help="sd: AutoencoderKL state-dict .pt (or 'random'): decode each candidate before scoring")
    parser.add_argument('--clip', type=str, default=None, help="sd, --scorer clip: CLIPModel state-dict .pt (or 'random')")
    parser.add_argument('--clip-text', dest='clip_text', type=str, default=None,
                        help='sd, --scorer clip: precomputed prompt embedding .pt [1, proj]')
    args = parser.parse_args()

    if args.backend == 'sd' and args.scorer == 'imagenet':
        raise ValueError('imagenet scorer is only available for edm backend')
    if args.backend == 'edm' and args.scorer == 'clip':
        raise ValueError('clip scorer is only available for sd backend')
    if args.backend == 'sd':
        return main_sd(args)

    scorer = get_scorer('edm', args.scorer, args.device, args.classifier)
    num_images = 1
    gridw = gridh = 1
    ddpmpp = args.network is None and args.arch == 'ddpmpp'
    res = 32 if ddpmpp else 64
    latents = torch.randn([num_images, 3, res, res])
    class_labels = None if ddpmpp else torch.eye(1000)[torch.randint(1000, size=[num_images])]
    device = torch.device(args.device)
    num_steps = 18

    from diffusion_tts_b200.edm.main import SamplingMethod, generate_image_grid
    method_map = {
        'naive': SamplingMethod.NAIVE,
        'rejection': SamplingMethod.REJECTION_SAMPLING,
        'beam': SamplingMethod.BEAM_SEARCH,
        'mcts': SamplingMethod.MCTS,
        'zero_order': SamplingMethod.ZERO_ORDER,
        'eps_greedy': SamplingMethod.EPS_GREEDY,
    }
    if args.method not in method_map:
        raise ValueError(f"Unknown method: {args.method}")
    sampling_params = {'scorer': scorer}
    if args.method in ['rejection', 'zero_order', 'eps_greedy', 'beam', 'mcts']:
        sampling_params.update(N=args.N, K=args.K, lambda_param=args.lambda_, eps=args.eps, B=args.B, S=args.S)
    network = args.network if args.network is not None else (random_init_ddpmpp() if ddpmpp else random_init_adm())
    outname = args.output or f"edm_{args.method}_{args.scorer}.png"
    generate_image_grid(network, outname, latents, class_labels, seed=args.seed, gridw=gridw, gridh=gridh,
                        device=device, num_steps=num_steps, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003,
                        sampling_method=method_map[args.method], sampling_params=sampling_params)
    print(f"\n[EDM] Saved: {outname}\n")


if __name__ == '__main__':
    main()

"""Command-line front end with the reference's flags (reference main.py:80-98), driving the
B200 noise-search path behind `--method/--scorer` for the EDM backend.

    python main.py --backend edm --scorer brightness --method eps_greedy --N 64 --K 20

Differences from the reference CLI, all forced by the offline B200 setting:
  * `--network` (new, optional): a local network pickle / `.pt` bundle.  The reference downloads
    the ImageNet-64 ADM pickle (main.py:157-158); without `--network` this front end builds a
    random-init ADM of the same architecture so that the path can run without network access.
  * `--classifier` (new, optional): a local `64x64_classifier.pt` for `--scorer imagenet` (random-init otherwise).
  * `--backend sd`, `--scorer clip` and `--method mcts` are not part of
    the B200 hot path yet and raise the same ValueError / NotImplementedError a wrong choice would.
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


def random_init_adm(seed=1234):
    """ImageNet-64 ADM bundle with every tensor drawn N(0, 1/fan_in) (the reference's zero-initialised
    conv1/proj/out_conv would make F_x == 0, SURVEY.md 7 hard part 2)."""
    from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
    return dict(state_dict=random_state_dict(adm_param_shapes(), seed), sigma_data=0.5)


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
    args = parser.parse_args()

    if args.backend == 'sd' and args.scorer == 'imagenet':
        raise ValueError('imagenet scorer is only available for edm backend')
    if args.backend == 'edm' and args.scorer == 'clip':
        raise ValueError('clip scorer is only available for sd backend')
    if args.backend == 'sd':
        raise NotImplementedError('the SD backend is not on the B200 hot path yet (SURVEY.md 8: a16, f1-f2)')

    scorer = get_scorer('edm', args.scorer, args.device, args.classifier)
    num_images = 1
    gridw = gridh = 1
    latents = torch.randn([num_images, 3, 64, 64])
    class_labels = torch.eye(1000)[torch.randint(1000, size=[num_images])]
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
    network = args.network if args.network is not None else random_init_adm()
    outname = args.output or f"edm_{args.method}_{args.scorer}.png"
    generate_image_grid(network, outname, latents, class_labels, seed=args.seed, gridw=gridw, gridh=gridh,
                        device=device, num_steps=num_steps, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003,
                        sampling_method=method_map[args.method], sampling_params=sampling_params)
    print(f"\n[EDM] Saved: {outname}\n")


if __name__ == '__main__':
    main()

"""Secondary benchmark (BASELINE.json configs[4]): SD-1.5-shaped 4x64x64 latent UNet2DConditionModel, random init,
--method beam with B=8 beams, brightness scorer on the Tweedie x0, candidates sharded over the GPUs of one box.
Weak scaling like bench.py: 32 candidates per GPU and step (N = 4 noises per beam per GPU; 8 GPUs = the named B=8 N=32).
One step = one DDIM timestep of the beam search: 1 UNet call at batch 2B (replicated), DDIM+CFG for all B*N candidates,
1 UNet call at batch 2*B*N/G on this rank's slice, x0 + quantise + score, all-gather of the scores, stable top-B.

  python tools/bench_sd_beam.py [--steps K] [--warmup W]          (or under torchrun, like bench.py)
Prints one JSON line in bench.py's format (metric: scored candidates/s)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--gpus', type=int, default=1)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--per-gpu', type=int, default=32)
ap.add_argument('--vae', action='store_true', help='decode every candidate x0 to a 512x512 image before scoring (SURVEY 8 f1)')
ap.add_argument('--clip', action='store_true', help='--vae plus the CLIP ViT-L/14 scorer (random-init) on every decoded image (SURVEY 8 f4)')
args = ap.parse_args()
args.vae = args.vae or args.clip
rank, world, lrank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lrank)
dev = torch.device('cuda', lrank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from bench import ClockSampler, peaks
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200 import ops
from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes
from diffusion_tts_b200.edm.main import Shard
from diffusion_tts_b200.sd.beam import DDIMTable, sd_beam_search
from diffusion_tts_b200.sd_unet import SDUNetEngine
from diffusion_tts_b200.arch import vae_decoder_param_shapes
from diffusion_tts_b200.vae import DecodedImageScorer, VAEDecoderEngine

B = 8
N = args.per_gpu * world // B
eng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device=dev)
g = torch.Generator().manual_seed(1)
eng.set_context(torch.randn(2, 77, 768, generator=g).to(dev))
lat = torch.randn(1, 4, 64, 64, generator=g).to(dev)
tab = DDIMTable(50)
shard = Shard(rank, world, None) if world > 1 else None
vae_kw = {}
if args.vae:
    veng = VAEDecoderEngine(random_state_dict(vae_decoder_param_shapes(), 4321), device=dev)
    score_fn = None
    if args.clip:
        from diffusion_tts_b200.arch import clip_param_shapes
        from diffusion_tts_b200.clip import CLIPScorer
        score_fn = CLIPScorer(random_state_dict(clip_param_shapes(vision_only=True), 33), device=dev)
        score_fn.set_text_embeds('a photo', torch.randn(1, 768, generator=g))
    vae_kw = dict(decode=lambda x0: x0, scorer=DecodedImageScorer(veng, score_fn, 'a photo', chunk=8))
total = args.warmup + args.steps
noises = {i: torch.randn(B, N, 4, 64, 64, generator=g).to(dev) for i in range(total)}      # same on every rank


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(step_ids):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    best, rec = sd_beam_search(eng, tab, lat, None, B, N, noises=noises, shard=shard, steps=step_ids, **vae_kw)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), rec


timed(list(range(args.warmup)))
clocks = ClockSampler(lrank)
clocks.start()
ops.LAUNCHES[0] = 0
ms, rec = timed(list(range(args.warmup, total)))
launches = ops.LAUNCHES[0]
clk = clocks.stop()
value = B * N * args.steps / (ms / 1e3)

fp = eng.plan(2 * args.per_gpu, 64)
fp.plan.run_timed()
per_op = fp.plan.run_timed()
gemm_ms = sum(t for t, k in zip(per_op, fp.plan.kinds) if k == 'gemm')
gemm_flops = sum(f for f, k in zip(fp.plan.flops, fp.plan.kinds) if k == 'gemm')
by_kind = {}
for t, k in zip(per_op, fp.plan.kinds):
    by_kind[k] = by_kind.get(k, 0.0) + t
peak_tf, _, _, peak_src = peaks()
if rank == 0:
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
    fwd_per_step = 2 * B + 2 * B * N // world                  # UNet forwards (samples) per rank and step
    print(json.dumps({
        'metric': 'scored_candidates_per_sec', 'value': value, 'unit': 'candidates/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if __import__('diffusion_tts_b200._lib', fromlist=['ACT_BF16']).ACT_BF16 else 'fp16', 'data': 'synthetic',
        'config': {'workload': f'SD-1.5-shaped UNet2DConditionModel (859.5M, random-init) 4x64x64 latents, beam B={B} N={N} '
                               f'(={args.per_gpu} candidates/GPU), DDIM eta=1 CFG 7.5, ' + ('SD-1.5 VAE decode (49.5M, random-init) of every Tweedie x0 to 512x512 + ' + ('CLIP ViT-L/14 scorer (random-init vision tower, Pillow-exact preprocessing)' if args.clip else 'RGB brightness') if args.vae else 'latent brightness on Tweedie x0'),
                   'B': B, 'N': N, 'candidates_per_gpu': args.per_gpu, 'unet_forwards_per_rank_step': fwd_per_step,
                   'l2': 'not flushed: 1.7 GB bf16 weights + ~9 GB activations per call exceed the 126 MB L2'},
        'gpu_launches': launches, 'clocks': clk,
        'roofline': {'bound': 'tensor', 'kernel': 'gemm_conv_kernel (tcgen05 implicit GEMM)', 'achieved': achieved,
                     'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'peak_source': peak_src,
                     'gemm_ms_per_forward_batch': gemm_ms, 'forward_ms': sum(per_op), 'ms_by_kernel_kind': by_kind,
                     'flops_per_forward_batch': gemm_flops,
                     'whole_step_tflops_gemm_only': fwd_per_step / (2 * args.per_gpu) * gemm_flops / (ms / args.steps / 1e3) / 1e12}}))
if world > 1:
    dist.destroy_process_group()

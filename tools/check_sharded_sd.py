"""Config 5 (BASELINE.json): SD beam search with the candidates (and the beams' own UNet call) sharded over the ranks of one
box.  Every rank runs the sharded search; rank 0 also runs it unsharded and checks that scores, kept indices and surviving
beams are BIT-IDENTICAL, then prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29514 \\
      tools/check_sharded_sd.py [--B 4] [--N 4] [--steps 2]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--B', type=int, default=4)
ap.add_argument('--N', type=int, default=4)
ap.add_argument('--steps', type=int, default=2)
args = ap.parse_args()
rank, world, lrank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lrank)
dev = torch.device('cuda', lrank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import random_state_dict, sd_unet_param_shapes
from diffusion_tts_b200.edm.main import Shard
from diffusion_tts_b200.sd.beam import DDIMTable, sd_beam_search
from diffusion_tts_b200.sd_unet import SDUNetEngine

eng = SDUNetEngine(random_state_dict(sd_unet_param_shapes(), 1234), device=dev)
g = torch.Generator().manual_seed(3)
ctx = torch.randn(2, 77, 768, generator=g).to(dev)
lat = torch.randn(1, 4, 64, 64, generator=g).to(dev)
tab = DDIMTable(50)
noises = {i: torch.randn(args.B, args.N, 4, 64, 64, generator=g).to(dev) for i in range(args.steps)}
steps = list(range(args.steps))
shard = Shard(rank, world, None) if world > 1 else None
best_s, rec_s = sd_beam_search(eng, tab, lat, ctx, args.B, args.N, noises=noises, shard=shard, record=True, steps=steps)
torch.cuda.synchronize()
ok = True
if rank == 0:
    best_1, rec_1 = sd_beam_search(eng, tab, lat, ctx, args.B, args.N, noises=noises, shard=None, record=True, steps=steps)
    torch.cuda.synchronize()
    same_scores = all(torch.equal(a, b) for a, b in zip(rec_s.scores, rec_1.scores))
    same_best = all(torch.equal(a, b) for a, b in zip(rec_s.best, rec_1.best))
    same_beams = all(torch.equal(a, b) for a, b in zip(rec_s.beams, rec_1.beams)) and torch.equal(best_s, best_1)
    ok = same_scores and same_best and same_beams
    print(json.dumps({'check': 'sd beam sharded == unsharded', 'B': args.B, 'N': args.N, 'world': world, 'steps': args.steps,
                      'scores_bit_identical': same_scores, 'kept_indices_equal': same_best, 'beams_bit_identical': same_beams,
                      'kept': [b.tolist() for b in rec_s.best]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)

"""Config 3 (BASELINE.json): ZERO_ORDER N=256 (local-neighbourhood perturbation, eps=0) with the candidates sharded
over the ranks of one box.  Every rank runs the sharded search; rank 0 also runs it unsharded and checks that
the selected indices and the committed trajectory are BIT-IDENTICAL, then prints one JSON line with the timing.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29512 \
      tools/check_sharded.py [--N 256] [--steps 4]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument('--N', type=int, default=256)
ap.add_argument('--steps', type=int, default=4)
ap.add_argument('--K', type=int, default=1)
args = ap.parse_args()
rank, world, lrank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lrank)
dev = torch.device('cuda', lrank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
from diffusion_tts_b200 import build
build.build()
from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
from diffusion_tts_b200.edm.main import SamplingParams, Shard, eps_greedy_search
from diffusion_tts_b200.scorers import BrightnessScorer

N = args.N
net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
table = StepTable(net, dev, 18, S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
g = torch.Generator().manual_seed(7)
labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
steps = [5, 9, 13, 16][:args.steps]
x0 = (torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) * 3).to(dev)
pre = {}
for i in steps:
    pre[i] = torch.randn(1, args.K, N, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
    pre[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).to(dev)
params = SamplingParams(N=N, K=args.K, eps=0.0, lambda_param=0.15, scorer=BrightnessScorer(device=dev))
shard = Shard(rank, world, None) if world > 1 else Shard()

def run(sh):
    torch.manual_seed(0)
    return eps_greedy_search(net, None, labels, params, table, precomputed_noise=pre, shard=sh, record=True,
                             step_indices=steps, x_init=x0)

run(shard)                                   # warm-up (plans, graphs, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
x_sh, rec_sh = run(shard)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = True
if rank == 0:
    x_1, rec_1 = run(Shard())                # unsharded on rank 0: all N candidates on one GPU
    idx_sh = torch.stack(rec_sh.indices).cpu()
    idx_1 = torch.stack(rec_1.indices).cpu()
    same_idx = bool((idx_sh == idx_1).all())
    same_x = all(torch.equal(a, b) for a, b in zip(rec_sh.x_steps, rec_1.x_steps)) and torch.equal(x_sh, x_1)
    ok = same_idx and same_x
    diffs = [float((a - b).abs().max()) for a, b in zip(rec_sh.x_steps, rec_1.x_steps)]
    pdiffs = [float((a - b).abs().max()) for a, b in zip(rec_sh.pivots, rec_1.pivots)]
    print(json.dumps({'check': 'zero_order sharded == unsharded', 'N': N, 'world': world, 'steps': steps,
                      'indices_sharded': idx_sh.flatten().tolist(), 'indices_unsharded': idx_1.flatten().tolist(),
                      'indices_equal': same_idx, 'trajectory_bit_identical': same_x, 'x_step_max_abs_diff': diffs, 'pivot_diff': pdiffs,
                      'ms_per_step_sharded': ms.item() / len(steps),
                      'candidates_per_sec': N * len(steps) / (ms.item() / 1e3)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)

"""Benchmark of the candidate-batched noise-search step (BASELINE.json metric):
scored candidates/sec, EDM ImageNet-64 ADM (DhariwalUNet, class-conditional, random init),
eps_greedy, brightness scorer, 18-step Heun schedule (S_churn=40, S_min=0.05, S_max=50,
S_noise=1.003) -- BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one denoising timestep of the eps_greedy search with K_local=1: build N=64
candidates per GPU around the pivot, 2 batched denoiser calls (1 on the last of 18 steps),
Tweedie x0, brightness score, first-max argmax (+ all-reduce of the packed key when N>1 GPUs),
pivot update and the commit step (2 more denoiser calls at batch 1).  Steps cycle through the
18 timesteps, so any multiple of 18 steps averages 426.5 GFLOP per scored candidate.

`value`  : candidates/s with all inputs (noise directions, pivots) resident in HBM.
`e2e`    : the same loop through the public API with HOST (pinned) noise buffers: per step the
           pivot and this rank's slice of the N direction tensors are copied host->device and the
           winning index, its score and the committed state are copied back into pinned host buffers
           (asynchronously, in stream order; the timed region ends with a synchronize, so every byte
           has arrived inside it).
`extras` : the same steps with the exact shortcut for the noise-free timesteps (the N identical
           candidates of such a step evaluated once) -- reported separately, never the headline.
`--impl reference` times the reference algorithm's CPU path (oracle port, all host threads) on
a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 64
NUM_STEPS = 18
SAMPLER = dict(S_churn=40, S_min=0.05, S_max=50, S_noise=1.003)
FLOP_PER_NFE = 219.33e9          # ADM-64 forward, per sample
METRIC = 'scored_candidates_per_sec'
UNIT = 'candidates/s'


def workload_name(n_total, scorer='brightness'):
    sc = 'brightness scorer' if scorer == 'brightness' else ('ImageNet classifier-probability scorer (EncoderUNetModel 64x64, '
                                                             'random-init, in the loop)' if scorer == 'imagenet' else
                                                             'compressibility scorer (exact JPEG q80 byte count)')
    return (f'EDM ImageNet-64 ADM (DhariwalUNet 295.9M, class-cond, random-init), eps_greedy N={n_total} '
            f'(={N_PER_GPU}/GPU) K=1 lambda=0.15 eps=0, {sc}, 18-step Heun cycle, b=1 image')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[j] for r in self.rows if len(r) >= 7 for j in range(4) if r[3 + j].lower() == 'active'})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1415.6), d.get('hbm_gbs', 6452.2), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------ reference arm / cpu baseline
def oracle_cpu_rate(n_cand: int, steps, warmup: int):
    """Reference algorithm on the host cores (oracle port of edm/main.py:714-860 + the fp32
    ADM-64 forward): `n_cand` candidates through each listed timestep, K=1."""
    from oracle import edm_oracle as O
    torch.manual_seed(0)
    spec = O.build_unet_spec('DhariwalUNet', 64, 3, 3, label_dim=1000)
    sd = O.seeded_state_dict(O.unet_param_shapes(spec), 1234)
    net = O.OracleNet(spec, sd)
    t_steps = O.karras_schedule(NUM_STEPS)
    g = torch.Generator().manual_seed(1)
    labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)]
    x = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) * t_steps[0]
    lam = 0.15 * (3 * 64 * 64) ** 0.5
    times = []
    for it, i in enumerate(steps):
        pivot = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64)
        dirs = [torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64) for _ in range(n_cand)]
        t0 = time.perf_counter()
        scales = [O.candidate_scale_fp32(((i * 31 + n * 17) % 1000) / 1000.0, lam) for n in range(n_cand)]
        cands = O.make_candidates(pivot, dirs, scales, [None] * n_cand)
        _, x0 = O.heun_step(net, x.repeat(n_cand, 1, 1, 1), t_steps[i], t_steps[i + 1], i, cands,
                            labels.repeat(n_cand, 1), num_steps=NUM_STEPS, **SAMPLER)
        scores = O.brightness_score(O.quantize_u8(x0)).reshape(n_cand, 1)
        best = O.argmax_first(scores, dim=0)
        pivot = cands[best[0]:best[0] + 1]
        x_new, _ = O.heun_step(net, x, t_steps[i], t_steps[i + 1], i, pivot, labels, num_steps=NUM_STEPS, **SAMPLER)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return n_cand * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_cand = 2
    steps = [(8 + j) % NUM_STEPS for j in range(args.warmup + args.steps)]
    rate, per_step = oracle_cpu_rate(n_cand, steps, args.warmup)
    cores = torch.get_num_threads()
    sample = (f'{n_cand} candidates per step (of the {N_PER_GPU} of the workload) through {args.steps} eps_greedy '
              f'timesteps incl. the commit step, fp32 ADM-64 on CPU')
    line = {'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(N_PER_GPU * args.gpus), 'sampled': sample},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build(oracle=False)
    from diffusion_tts_b200 import ops
    from diffusion_tts_b200.arch import adm_param_shapes, random_state_dict
    from diffusion_tts_b200.denoiser import B200Denoiser, StepTable
    from diffusion_tts_b200.edm.main import SamplingParams, Shard, eps_greedy_search
    from diffusion_tts_b200.scorers import BrightnessScorer

    N = N_PER_GPU * world
    net = B200Denoiser(random_state_dict(adm_param_shapes(), 1234), device=dev)
    table = StepTable(net, dev, NUM_STEPS, **SAMPLER)
    shard = Shard(rank, world, None)
    if args.scorer == 'imagenet':          # BASELINE.json configs[3]: the classifier network runs inside the search loop
        from diffusion_tts_b200.arch import classifier_param_shapes
        from diffusion_tts_b200.classifier import ImageNetScorer
        scorer = ImageNetScorer(random_state_dict(classifier_param_shapes(), 22), device=dev)
    elif args.scorer == 'compressibility':
        from diffusion_tts_b200.scorers import CompressibilityScorer
        scorer = CompressibilityScorer(device=dev)
    else:
        scorer = BrightnessScorer(device=dev)
    params = SamplingParams(N=N, K=1, eps=0.0, lambda_param=0.15, scorer=scorer)
    g = torch.Generator().manual_seed(1)
    latents = torch.randn(1, 3, 64, 64, generator=g)
    labels = torch.eye(1000)[torch.randint(1000, (1,), generator=g)].to(dev)
    total = args.warmup + args.steps
    order = [j % NUM_STEPS for j in range(total)]
    # synthetic noise of the named shapes; identical on every rank (same seed), as a shared RNG stream would be
    host = {}
    for i in sorted(set(order)):
        host[f'pivot_{i}'] = torch.randn(1, 3, 64, 64, generator=g, dtype=torch.float64).pin_memory()
        host[i] = torch.randn(1, 1, N, 3, 64, 64, generator=g, dtype=torch.float64).pin_memory()
    on_dev = {k: v.to(dev) for k, v in host.items()}
    x0 = (latents.to(torch.float64) * table.t_steps[0].cpu()).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    esc = None if args.escalate < 0 else bool(args.escalate)

    def timed(noise, steps_idx, x_init, on_step=None, dedupe=False, escalate=esc):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, rec = eps_greedy_search(net, None, labels, params, table, precomputed_noise=noise, shard=shard,
                                   step_indices=steps_idx, x_init=x_init, on_step=on_step, prefetch=bool(args.prefetch),
                                   dedupe_noise_free=dedupe, escalate=escalate, **({'kappa': args.kappa} if args.kappa > 0 else {}))
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), x, rec

    # ---- device-resident run
    _, x, _ = timed(on_dev, order[:args.warmup], x0)
    clocks = ClockSampler(local_rank)
    clocks.start()                      # every rank samples its own GPU; rank 0's goes into `clocks`, all into `per_rank`
    ops.LAUNCHES[0] = 0
    ms, x, rec = timed(on_dev, order[args.warmup:], x)
    launches = ops.LAUNCHES[0]
    clk = clocks.stop()
    value = N * args.steps / (ms / 1e3)

    # ---- reported separately, NOT the headline: the same steps with the exact shortcut for the noise-free timesteps (all N
    # candidates of such a step are one tensor: evaluate it once); results are bit-identical
    timed(on_dev, order[:args.warmup], x0, dedupe=True)
    ms_dd, x_dd, _ = timed(on_dev, order[args.warmup:], x, dedupe=True)
    noise_free = sum(1 for j in order[args.warmup:] if table.steps[j].s == 0.0)

    # ---- end-to-end run: host (pinned) noise in, per-step results out
    results = []

    # per-step results land in pinned host buffers (asynchronous device->host copies in stream order, like the
    # host->device noise copies); the timed region ends with a synchronize, so every byte has arrived inside it
    pinned = [(torch.empty(1, dtype=torch.int64).pin_memory(), torch.empty(1, dtype=torch.float32).pin_memory(),
               torch.empty(x0.shape, dtype=torch.float64).pin_memory()) for _ in range(total)]

    def read_back(i, x_next, idx, scores):
        if not args.async_readback:
            results.append((idx.cpu(), scores.max().cpu(), x_next.cpu()))
            return
        bi, bs, bx = pinned[len(results) % total]
        bi.copy_(idx, non_blocking=True)
        bs.copy_(scores.max().reshape(1), non_blocking=True)
        bx.copy_(x_next, non_blocking=True)
        results.append((bi, bs, bx))

    _, xe, _ = timed(host, order[:args.warmup], x0, read_back)
    ms_e2e, xe, _ = timed(host, order[args.warmup:], xe, read_back)
    e2e = N * args.steps / (ms_e2e / 1e3)
    # every rank uploads the step's pivot noise and ITS slice of the N candidate directions (fp64)
    h2d_rank = host[f'pivot_{order[-1]}'].numel() * 8 + host[order[-1]].numel() * 8 // world
    h2d = h2d_rank * world
    d2h = 8 + 4 + x0.numel() * 8

    # ---- dominant kernel (tcgen05 implicit-GEMM conv) roofline: per-op CUDA-event timing of one NFE at B = N/GPU
    fp = net.engine.plan(N_PER_GPU, 1)
    fp.plan.run_timed()
    per_op = fp.plan.run_timed()
    gemm_ms = sum(t for t, k in zip(per_op, fp.plan.kinds) if k == 'gemm')
    gemm_flops = sum(f for f, k in zip(fp.plan.flops, fp.plan.kinds) if k == 'gemm')
    n_gemm = sum(1 for k in fp.plan.kinds if k == 'gemm')
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
    by_kind = {}
    for t, k in zip(per_op, fp.plan.kinds):
        by_kind[k] = by_kind.get(k, 0.0) + t
    nfe_ms = sum(per_op)

    per_rank = None
    if world > 1:                       # which GPU is the slow one?  (weak scaling waits for the slowest rank every step)
        mine = {'rank': rank, 'sm_mhz': clk.get('sm_mhz'), 'reasons': clk.get('reasons'), 'nfe_ms': nfe_ms}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = {'sm_mhz': [g['sm_mhz'] for g in gathered], 'nfe_ms': [round(g['nfe_ms'], 3) for g in gathered],
                    'reasons': sorted({r for g in gathered for r in (g['reasons'] or [])})}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, per_step = oracle_cpu_rate(2, [8, 9, 10], 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '2 candidates x 2 timed eps_greedy timesteps (i=9,10; 1 warm-up) incl. commit, fp32 ADM-64 '
                         'oracle port on the host cores'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': workload_name(N, args.scorer), 'N_per_gpu': N_PER_GPU, 'K': 1, 'num_steps': NUM_STEPS,
                   'l2': 'not flushed: per-step working set (0.6 GB bf16 weights + >2 GB activations) exceeds the 126 MB L2',
                   'sampler_state': 'fp64', 'unet': 'bf16 storage, fp32 accumulate/GroupNorm/softmax'},
        'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h * world,
                'h2d_bytes_per_step_per_rank': h2d_rank, 'd2h_bytes_per_step_per_rank': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'escalation': {'mode': args.escalate, 'rows_refined_per_step': rec.escalated, 'truncated_rounds': rec.truncated},
        'extras': {'value_with_noise_free_dedupe': N * args.steps / (ms_dd / 1e3), 'ms_per_step': ms_dd / args.steps,
                   'noise_free_steps': noise_free,
                   'note': 'same candidates counted; on the timesteps with noise scale 0 the N identical candidates are '
                           'evaluated once (bit-identical results, eps_greedy_search(dedupe_noise_free=True)); not the headline'},
        'clocks': clk,
        'per_rank': per_rank,
        'roofline': {'bound': 'tensor', 'kernel': 'gemm_conv_kernel (tcgen05 implicit-GEMM conv3x3/1x1)',
                     'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                     'peak_source': peak_src, 'traffic': 591.9e6,
                     'traffic_note': 'bytes; ncu dram read+write of the largest of the GEMM launches (dec.64x64_up.conv1, '
                                     '540 us of the NFE) vs 604e6 algorithmic (A + residual + out + weights): '
                                     'profiles/r01_ncu_full_v3_summary.txt',
                     'launches_per_nfe': n_gemm,
                     'flops_per_nfe_batch': gemm_flops, 'gemm_ms_per_nfe': gemm_ms, 'nfe_ms': nfe_ms,
                     'ms_by_kernel_kind': by_kind,
                     'whole_step_tflops': (N_PER_GPU * 35 / 18 + 35 / 18) * FLOP_PER_NFE / (ms / args.steps / 1e3) / 1e12},
        'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=18)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--kappa', type=float, default=0.0, help='escalation threshold in units of the score spread (0 = API default)')
    ap.add_argument('--escalate', type=int, default=-1, help='near-tie precision escalation: 1 on, 0 off, -1 = API default (on)')
    ap.add_argument('--prefetch', type=int, default=0, help='e2e: stage the next step\'s host noise on a side stream (no measurable gain)')
    ap.add_argument('--async-readback', type=int, default=1, help='e2e: per-step results into pinned buffers, asynchronously')
    ap.add_argument('--scorer', type=str, default='brightness', choices=['brightness', 'imagenet', 'compressibility'],
                    help='brightness = BASELINE.json configs[1] (the headline); imagenet = configs[3]')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()

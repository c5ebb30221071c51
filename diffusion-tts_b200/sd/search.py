"""eps_greedy / zero_order / naive sampling of the SD backend on B200.

Semantics = the reference's `StableDiffusionPipeline.__call__(method=..., params={"N","K","lambda","eps"})` default branch,
sd/diffusers/src/diffusers/pipelines/stable_diffusion/pipeline_stable_diffusion.py:1330-1437, with the reference's edited
DDIM step (scheduling_ddim.py:342-471, eta = 1, variance noise supplied):

  per timestep t:   eps = CFG(UNet([x; x], t));  pivot = randn_like(x)                                   (:1341-1366)
    K rounds:       N candidates: w.p. eps (eps_greedy only) a fresh N(0,I) draw, else
                    pivot + randn/||randn|| * rand * lambda * sqrt(C*H*W)                                 (:1371-1379)
                    for each: x_c = DDIM(eps, t, x, variance_noise = candidate)                           (:1384)
                              eps2 = CFG(UNet([x_c; x_c], t))   -- the SAME t                             (:1392-1406)
                              score(decode(pred_original_sample(eps2, t, x_c)))                           (:1412-1431)
                    pivot = first best-scoring candidate                                                  (:1434-1435)
    x = DDIM(eps, t, x, variance_noise = pivot)                                                           (:1437)

Execution: the N candidates of a round go through ONE UNet call of batch 2N (both CFG halves; the prompt K/V were
projected once per prompt), candidate construction / DDIM+CFG / x0+quantise+score are one kernel each, the argmax stays on
the device.  RNG: the reference draws `torch.rand(1)` on the CPU generator and `torch.randn_like(latents)` on the device
generator; both sequences are reproduced call for call (the CPU draws cost no device synchronisation).
Candidates can be sharded over a process group exactly as in the beam search (scores all-gathered)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from .. import ops
from ..sd_unet import SDUNetEngine
from .beam import DDIMTable


@dataclass
class SDSearchRecord:
    scores: List[torch.Tensor] = field(default_factory=list)     # per round: [N]
    best: List[torch.Tensor] = field(default_factory=list)       # per round: [1] index of the selected candidate
    cands: List[torch.Tensor] = field(default_factory=list)      # per round: [N,4,H,W] candidate noises
    x: List[torch.Tensor] = field(default_factory=list)          # per step: committed latents
    max_score: Optional[torch.Tensor] = None
    scored_candidates: int = 0


@torch.no_grad()
def sd_eps_greedy_search(eng: SDUNetEngine, table: DDIMTable, latents: torch.Tensor, ctx_pair: Optional[torch.Tensor], N: int,
                         K: int, lam: float, eps: float, method: str = 'eps_greedy', *, guidance_scale: float = 7.5,
                         noise: Optional[dict] = None, decode: Optional[Callable] = None, scorer: Optional[Callable] = None,
                         shard=None, record: bool = False, teacher_x: Optional[List[torch.Tensor]] = None,
                         steps: Optional[List[int]] = None):
    """latents [1,4,H,W]; `noise` (optional, parity tests) as in oracle/sd_oracle.py:eps_greedy_search.
    Returns (final latents [1,4,H,W], SDSearchRecord)."""
    if method not in ('eps_greedy', 'zero_order', 'naive'):
        raise ValueError(f"Unknown method: {method}")
    if (decode is None) != (scorer is None):
        raise ValueError('decode and scorer go together: give both (generic path) or neither (fused latent brightness)')
    dev = eng.device
    if ctx_pair is not None:
        eng.set_context(ctx_pair)
    x = latents.to(device=dev, dtype=torch.float32).contiguous()
    if x.shape[0] != 1:
        raise ValueError('the SD search runs one prompt at a time, like the reference pipeline')
    C, H, W = x.shape[1:]
    E = C * H * W
    sqrt_e = float(np.sqrt(W * H * C))
    world = shard.world if shard is not None else 1
    lo, hi = shard.bounds(N) if shard is not None else (0, N)
    fp1 = eng.plan(2, H)
    fp2 = eng.plan(2 * (hi - lo), H) if method != 'naive' else None
    rec = SDSearchRecord()
    last_best_score = None

    def score_rows(eps2, rows, cf):
        if scorer is None:
            return ops.ddim_x0_score(eps2, rows, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'])[0]
        _, _, x0 = ops.ddim_x0_score(eps2, rows, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'], want_x0=True)
        return torch.as_tensor(scorer(decode(x0))).to(device=dev, dtype=torch.float32).reshape(-1)

    for i in (range(len(table.timesteps)) if steps is None else steps):
        t = table.timesteps[i]
        cf = table.coeffs(t)
        fp1.x_in[:1].copy_(x)
        fp1.x_in[1:].copy_(x)
        eps1 = eng.run(fp1, t)
        pivot = (noise['pivot'][i].to(device=dev, dtype=torch.float32).contiguous() if noise is not None
                 else torch.randn_like(x))                                          # :1366
        if method != 'naive':
            for k in range(K):
                # ---- the reference's draws, call for call (:1371-1379): CPU rand(1) for the branch and the scale,
                # device randn_like for the direction / the fresh noise
                fresh = np.zeros(N, dtype=np.uint8)
                u = np.zeros(N, dtype=np.float32)
                dirs = []
                for n in range(N):
                    r = float(noise['r'][i][k][n]) if noise is not None else torch.rand(1).item()
                    if (r < eps) if method == 'eps_greedy' else 0.0:
                        fresh[n] = 1
                    if noise is None:
                        dirs.append(torch.randn_like(x))
                    if not fresh[n]:
                        u[n] = float(noise['u'][i][k][n]) if noise is not None else torch.rand(1).item()
                if noise is None:
                    # one DISCARDED device draw per scored candidate: the reference's scoring-only second scheduler.step runs
                    # with eta = 1 and no variance_noise (:1412 -> scheduling_ddim.py:457-461); without it every later pivot /
                    # direction of a free-running run would differ from the reference at the same seed
                    for _ in range(N):
                        torch.randn_like(x)
                D = (noise['dirs'][i][k].to(device=dev, dtype=torch.float32).contiguous() if noise is not None
                     else torch.cat(dirs))
                cands = ops.sd_candidates(pivot, D, torch.from_numpy(u).to(dev), lam, sqrt_e,
                                          torch.from_numpy(fresh).to(dev) if fresh.any() else None)
                lat_c = ops.ddim_cfg_step(eps1, x, cands, N, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'],
                                          cf['sqrt_alpha_prev'], cf['dir_coef'], cf['std'])
                local = lat_c[lo:hi]
                fp2.x_in[:hi - lo].copy_(local)
                fp2.x_in[hi - lo:].copy_(local)
                eps2 = eng.run(fp2, t)
                scores = score_rows(eps2, local, cf)
                rec.scored_candidates += hi - lo
                if world > 1:
                    import torch.distributed as dist
                    allv = torch.empty(N, device=dev, dtype=torch.float32)
                    dist.all_gather_into_tensor(allv, scores.contiguous(), group=shard.group)
                    scores = allv
                best = ops.argmax_first(scores.reshape(N, 1).contiguous())          # first maximal candidate (:1434-1435)
                pivot = cands.index_select(0, best)
                last_best_score = scores.index_select(0, best)
                if record:
                    rec.scores.append(scores)
                    rec.best.append(best)
                    rec.cands.append(cands)
        x = ops.ddim_cfg_step(eps1, x, pivot, 1, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'], cf['sqrt_alpha_prev'],
                              cf['dir_coef'], cf['std'])                             # :1437
        if record:
            rec.x.append(x)
        if teacher_x is not None:
            x = teacher_x[i].to(device=dev, dtype=torch.float32).contiguous()
    if last_best_score is None:                                                     # naive: score the result (:1469-1474)
        if scorer is None:
            zero = torch.zeros(2, H, W, C, device=dev, dtype=torch.float32)
            last_best_score = ops.ddim_x0_score(zero, x, 0.0, 0.0, 1.0)[0]
        else:
            last_best_score = torch.as_tensor(scorer(decode(x))).to(device=dev, dtype=torch.float32).reshape(-1)
    rec.max_score = last_best_score.reshape(())
    return x, rec


@torch.no_grad()
def sd_mcts_as_shipped(eng: SDUNetEngine, table: DDIMTable, latents: torch.Tensor, ctx_pair: Optional[torch.Tensor], N: int,
                       S: int, *, guidance_scale: float = 7.5, decode: Optional[Callable] = None,
                       scorer: Optional[Callable] = None, record: bool = False):
    """`method="mcts"` of the SD backend, AS SHIPPED (pipeline_stable_diffusion.py:1172-1333).  The reference's tree never
    receives a reward: no rollout is scored, `visits` / `total_reward` are never updated, so the selection walk stays at the
    root, the first min(N, S) iterations each add one child (a DDIM step with a fresh `randn_like`), every iteration runs a
    discarded rollout, and `max(children, key = -inf)` returns the FIRST child.  The observable behaviour -- pinned against the
    real pipeline by tests/test_sd_oracle.py::test_mcts_as_shipped_matches_the_real_pipeline -- is a DDIM step with the step's
    first noise draw, plus the RNG consumption of the discarded work, reproduced here draw for draw (one child draw per
    expansion, one per rollout step: scheduler.step with eta = 1 and no variance_noise, scheduling_ddim.py:457).  The UNet
    evaluations whose results the reference throws away are not executed: S * (T - i + 2) network calls per step become 1."""
    if (decode is None) != (scorer is None):
        raise ValueError('decode and scorer go together: give both (generic path) or neither (fused latent brightness)')
    dev = eng.device
    if ctx_pair is not None:
        eng.set_context(ctx_pair)
    x = latents.to(device=dev, dtype=torch.float32).contiguous()
    C, H, W = x.shape[1:]
    fp1 = eng.plan(2, H)
    rec = SDSearchRecord()
    T = len(table.timesteps)
    for i, t in enumerate(table.timesteps):
        first = None
        for s in range(S):
            if s < N:                                                    # expansion: one child per iteration (:1216-1251)
                noise = torch.randn_like(x)
                first = noise if first is None else first
            for _ in range(i, T):                                        # the discarded rollout's draws (:1276-1300)
                torch.randn_like(x)
        if first is not None:                                            # "best" child = the first one (:1305-1308)
            cf = table.coeffs(t)
            fp1.x_in[:1].copy_(x)
            fp1.x_in[1:].copy_(x)
            eps1 = eng.run(fp1, t)
            x = ops.ddim_cfg_step(eps1, x, first, 1, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'], cf['sqrt_alpha_prev'],
                                  cf['dir_coef'], cf['std'])
        if record:
            rec.x.append(x)
    if scorer is None:                                                   # max_score is None -> the final result is scored (:1466)
        zero = torch.zeros(2, H, W, C, device=dev, dtype=torch.float32)
        score = ops.ddim_x0_score(zero, x, 0.0, 0.0, 1.0)[0]
    else:
        score = torch.as_tensor(scorer(decode(x))).to(device=dev, dtype=torch.float32).reshape(-1)
    rec.max_score = score.reshape(())
    return x, rec

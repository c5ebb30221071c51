"""Beam search of the SD backend (BASELINE.json config 5) on B200.

Semantics = the reference's `StableDiffusionPipeline.__call__(method="beam", params={"B":..,"N":..})`,
sd/diffusers/src/diffusers/pipelines/stable_diffusion/pipeline_stable_diffusion.py:1045-1170, with the reference's
edited DDIM step (scheduling_ddim.py:342-471: eta = 1, supplied variance noise, returns pred_original_sample):

  per timestep t, for every beam:  eps = CFG(UNet([x; x], t))                                   (:1058-1075)
    for N fresh noises:            cand = DDIM(eps, t, x, variance_noise)                         (:1080-1083)
                                   eps2 = CFG(UNet([cand; cand], t))   -- the SAME t              (:1086-1103)
                                   x0   = pred_original_sample(eps2, t, cand)                      (:1109)
                                   score(decode(x0))                                               (:1111-1123)
  keep the B best of the B*N candidates, stable sort descending = lowest flat index wins ties      (:1132-1134)
  final answer = the best-scoring surviving beam                                                    (:1153-1166)

What changes is the execution: the reference runs B*(1+N) batch-2 UNet calls per step in Python loops with a host
sync per candidate (:1123); here one UNet call of batch 2B and ONE of batch 2*B*N (all candidates, both CFG halves
share the two prompt contexts whose cross-attention K/V were projected once), the DDIM / guidance / Tweedie-x0 /
quantise / score arithmetic in two fused kernels, the top-B on the device, no host sync in the loop.
`decode` is injectable: identity (latent-space scoring, config 5) is fused; anything else falls back to
`scorer(decode(x0))` on materialised tensors.  Candidates can be sharded over the ranks of a process group: every
rank builds all candidate latents (cheap), runs the big UNet call on its slice only, and the scores are all-gathered;
the beams' own UNet call is split over the ranks as well (eps all-gathered) when the rank count divides B.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from .. import ops
from ..sd_unet import SDUNetEngine


class DDIMTable:
    """DDIMScheduler(beta_start=.00085, beta_end=.012, 'scaled_linear', clip_sample=False, set_alpha_to_one=False,
    steps_offset=1).set_timesteps(n): scheduling_ddim.py:190-216 (betas, alphas_cumprod), :297-340 ('leading')."""

    def __init__(self, num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, steps_offset: int = 1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps, self.num_inference_steps = num_train_timesteps, num_inference_steps
        ratio = num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset
        self.timesteps = [int(v) for v in ts]

    def coeffs(self, t: int, eta: float = 1.0) -> dict:
        """fp32 scalars of one step, formed with the same torch expressions as the reference (:398-440)."""
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        variance = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        return dict(sqrt_beta_t=float((1 - a_t) ** 0.5), sqrt_alpha_t=float(a_t ** 0.5), sqrt_alpha_prev=float(a_prev ** 0.5),
                    dir_coef=float((1 - a_prev - std ** 2) ** 0.5), std=float(std))


@dataclass
class BeamRecord:
    scores: List[torch.Tensor] = field(default_factory=list)     # per step: [B*N] (global candidate order: beam-major)
    best: List[torch.Tensor] = field(default_factory=list)       # per step: [B] flat indices of the kept candidates
    beams: List[torch.Tensor] = field(default_factory=list)      # per step: [B,4,H,W] surviving latents
    final_score: Optional[torch.Tensor] = None
    scored_candidates: int = 0


@torch.no_grad()
def sd_beam_search(eng: SDUNetEngine, table: DDIMTable, latents: torch.Tensor, ctx_pair: Optional[torch.Tensor], B: int,
                   N: int, *, guidance_scale: float = 7.5, noises: Optional[List[torch.Tensor]] = None,
                   decode: Optional[Callable] = None, scorer: Optional[Callable] = None, shard=None, record: bool = False,
                   teacher_beams: Optional[List[torch.Tensor]] = None, steps: Optional[List[int]] = None):
    """latents [1,4,H,W]; ctx_pair [2,T,D] = [negative/uncond, prompt] embeddings (None: keep the engine's context);
    noises[i] (optional) = [B, N, 4, H, W] variance noise of step i (drawn with torch.randn like :1080 otherwise).
    Returns (best latent [1,4,H,W], BeamRecord)."""
    dev = eng.device
    if ctx_pair is not None:
        eng.set_context(ctx_pair)
    if (decode is None) != (scorer is None):
        raise ValueError('decode and scorer go together: give both (generic path) or neither (fused latent brightness)')
    x = latents.to(device=dev, dtype=torch.float32)
    if x.shape[0] != 1:
        raise ValueError('beam search runs one prompt at a time, like the reference pipeline')
    C, H, W = x.shape[1:]
    R = B * N
    world = shard.world if shard is not None else 1
    lo, hi = shard.bounds(R) if shard is not None else (0, R)
    beams = x.expand(B, -1, -1, -1).contiguous()                       # :1046: B copies of the initial latents
    # the beams' own UNet call is sharded too when the ranks divide B: rank r evaluates beams [r*B/G, (r+1)*B/G) and the
    # eps halves are all-gathered (64 KB per beam); the engine is batch-invariant, so every rank ends with the same bits
    shard_beams = world > 1 and B % world == 0
    Bp = B // world if shard_beams else B
    b_lo = shard.rank * Bp if shard_beams else 0
    fp1 = eng.plan(2 * Bp, H)
    fp2 = eng.plan(2 * (hi - lo), H)
    rec = BeamRecord()
    step_ids = list(range(len(table.timesteps))) if steps is None else steps
    for i in step_ids:
        t = table.timesteps[i]
        cf = table.coeffs(t)
        # ---- eps of the beams: one UNet call of batch 2B
        mine = beams[b_lo:b_lo + Bp]
        fp1.x_in[:Bp].copy_(mine)
        fp1.x_in[Bp:].copy_(mine)
        eps1 = eng.run(fp1, t)
        if shard_beams:
            import torch.distributed as dist
            allg = torch.empty((world,) + tuple(eps1.shape), device=dev, dtype=eps1.dtype)
            dist.all_gather_into_tensor(allg, eps1.contiguous(), group=shard.group)
            eps1 = torch.cat([allg[:, :Bp].reshape(B, *eps1.shape[1:]), allg[:, Bp:].reshape(B, *eps1.shape[1:])])
        # ---- all B*N candidates (DDIM step with per-candidate variance noise)
        if noises is not None:
            nz = noises[i].to(device=dev, dtype=torch.float32).reshape(R, C, H, W).contiguous()
        else:
            # the reference's draws, call for call (oracle/sd_oracle.py:draw_beam_noise, pinned against the real pipeline):
            # per beam N separate randn_like([1,C,H,W]) (:1080), then one DISCARDED draw per scored candidate -- the
            # scoring-only second scheduler.step runs with eta = 1 and no variance_noise (:1109 -> scheduling_ddim.py:457)
            per = []
            for _b in range(B):
                per.append(torch.cat([torch.randn(1, C, H, W, device=dev) for _ in range(N)]))
                for _ in range(N):
                    torch.randn(1, C, H, W, device=dev)
            nz = torch.stack(per).reshape(R, C, H, W)
        cand = ops.ddim_cfg_step(eps1, beams, nz, N, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'],
                                 cf['sqrt_alpha_prev'], cf['dir_coef'], cf['std'])
        # ---- second UNet call at the same t on this rank's slice, both CFG halves
        local = cand[lo:hi]
        fp2.x_in[:hi - lo].copy_(local)
        fp2.x_in[hi - lo:].copy_(local)
        eps2 = eng.run(fp2, t)
        if scorer is None:
            scores, _, _ = ops.ddim_x0_score(eps2, local, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'])
        else:
            _, _, x0 = ops.ddim_x0_score(eps2, local, guidance_scale, cf['sqrt_beta_t'], cf['sqrt_alpha_t'], want_x0=True)
            scores = torch.as_tensor(scorer(decode(x0))).to(device=dev, dtype=torch.float32).reshape(-1)
        rec.scored_candidates += hi - lo
        if world > 1:
            import torch.distributed as dist
            allv = torch.empty(R, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(allv, scores.contiguous(), group=shard.group)
            scores = allv
        # ---- stable top-B: descending, lowest flat index first among equals (:1132-1134)
        order = torch.sort(scores, descending=True, stable=True).indices[:B]
        beams = cand.index_select(0, order).contiguous()
        if record:
            rec.scores.append(scores)
            rec.best.append(order)
            rec.beams.append(beams)
        if teacher_beams is not None:
            beams = teacher_beams[i].to(device=dev, dtype=torch.float32).contiguous()
    # ---- final pick (:1153-1166): score the surviving beams themselves; first strict maximum wins
    if scorer is None:
        zero = torch.zeros(2 * B, H, W, C, device=dev, dtype=torch.float32)
        fs, _, _ = ops.ddim_x0_score(zero, beams, 0.0, 0.0, 1.0)
    else:
        fs = torch.as_tensor(scorer(decode(beams))).to(device=dev, dtype=torch.float32).reshape(-1)
    best = ops.argmax_first(fs.reshape(B, 1).contiguous())
    rec.final_score = fs.max()
    return beams.index_select(0, best), rec

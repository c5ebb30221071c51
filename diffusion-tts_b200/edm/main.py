"""Drop-in for the reference's `edm/main.py` search driver on B200.

Same names and signatures as the reference module (edm/main.py:27-55): `SamplingMethod`,
`SamplingParams`, `generate_image_grid(...)`; same `precomputed_noise` protocol
(edm/main.py:114-117, 734-737, 753-755, 791-792) and the same torch RNG call sequence, so a
run without precomputed noise draws exactly the numbers the reference would draw on the same
device and seed.  What changes is everything underneath: the candidate fan-out, the two
batched denoiser calls, Tweedie x0, scoring and the argmax all run as hand-written sm_100a
kernels (libb200ns.so), the loop never synchronises with the host, and the candidate
dimension can be sharded over the ranks of a torch.distributed (NCCL) process group.

Deviations from the reference, each deliberate (SURVEY.md hard part 11):
  * no `os.environ["CUDA_VISIBLE_DEVICES"] = "0"` at import (edm/main.py:12) -- it would put
    every torchrun rank on GPU 0;
  * BEAM_SEARCH implements the evident intent (k = B beams, b = N noises per beam, stable
    top-k with lowest-index tie rule); the reference branch raises AttributeError at :140;
  * MCTS (`mcts_search`) keeps the reference's tree logic and RNG streams but batches every expansion and the rollouts
    of a simulation group.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from enum import Enum, auto
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import ops, philox
from .._lib import ACT_BF16
from ..denoiser import B200Denoiser, HeunStepper, StepTable, load_network
from ..scorers import BrightnessScorer, Scorer


class SamplingMethod(Enum):
    MCTS = auto()
    BEAM_SEARCH = auto()
    ZERO_ORDER = auto()
    NAIVE = auto()
    REJECTION_SAMPLING = auto()
    EPS_GREEDY = auto()


@dataclass
class SamplingParams:
    B: int = 2
    N: int = 4
    K: int = 20
    lambda_param: float = 0.15
    eps: float = 0.4
    S: int = 8
    scorer: Scorer = field(default_factory=lambda: BrightnessScorer(dtype=torch.float32))


@dataclass
class SearchRecord:
    """Optional trace of a run (device tensors; nothing here forces a sync)."""
    scores: List[torch.Tensor] = field(default_factory=list)      # per round: [N, b] (global N)
    indices: List[torch.Tensor] = field(default_factory=list)     # per round: [b] global candidate index
    pivots: List[torch.Tensor] = field(default_factory=list)      # per step: committed noise
    x_steps: List[torch.Tensor] = field(default_factory=list)     # per step: committed x_next (fp64)
    final_image: Optional[torch.Tensor] = None
    final_scores: Optional[torch.Tensor] = None
    scored_candidates: int = 0
    # near-tie escalation (always filled: host integers, no sync of their own)
    escalated: List[int] = field(default_factory=list)            # per round: rows re-scored by the precise engine (this rank)
    refined: List[Optional[torch.Tensor]] = field(default_factory=list)   # per round (record only): [N_local, b] refined scores, -inf = not a contender
    truncated: int = 0                                            # rounds whose contender list exceeded max_contenders
    mispredicted: int = 0                                         # speculated rounds whose refined winner differed (rolled back)
    escalation_log: List[tuple] = field(default_factory=list)     # per escalated round: (timestep, winner's lead / std per image, speculated?)
    missed_steps: List[int] = field(default_factory=list)         # timesteps of the rolled-back rounds


# Near-tie escalation (SURVEY.md 7 hard part 1b).  A candidate is a contender when its 16-bit score is within
# delta = KAPPA x std_n(scores) of the round's best; contenders are re-scored by the fp32-faithful engine (precise.py) and
# the argmax is taken over the refined scores.  What matters is the part of the 16-bit score error that DIFFERS between
# candidates (the part common to all candidates of a round, 2e-5, cancels in the comparison): measured against the real
# reference on the ADM-64 N=64 fixtures it has a standard deviation of 0.054-0.082 of the spread of the scores themselves at
# every noise level (fp16 storage; 0.17-0.26 with bf16), i.e. the difference of two candidates' errors has sigma ~0.1 spread:
# KAPPA = 0.35 is ~3.5 sigma (tests/test_full_parity_gpu.py dumps the tables).
# The bfloat16-storage build (B200NS_ACT=bf16) carries 3-4x the noise (its worst observed gain of a candidate on the true best
# is 0.5 of the spread: a flip at kappa = 0.35 on the K = 2 fixture), so its defaults are wider.
ESCALATION_KAPPA = 1.25 if ACT_BF16 else 0.35
MAX_CONTENDERS = 16 if ACT_BF16 else 8


@dataclass
class Shard:
    """Candidate sharding over a process group: rank r owns n in [r*N/G, (r+1)*N/G)."""
    rank: int = 0
    world: int = 1
    group: Any = None

    def bounds(self, N: int):
        if N % self.world:
            raise ValueError(f'N={N} must be divisible by the number of ranks ({self.world})')
        per = N // self.world
        return self.rank * per, (self.rank + 1) * per


def _score_rows(scorer, stepper: HeunStepper, x_cur, eps, i, labels_rows, C, HW, want_x_next=False):
    """Evaluate `eps` rows and score their Tweedie x0 -> (scores [R], x_next or None).
    Fused path for scorers that take channel sums (no uint8 image is materialised)."""
    if getattr(scorer, 'fused_sums', False):
        x_next, _, sums = stepper.step(x_cur, eps, i, want_x_next=want_x_next, want_sums=True)
        return scorer.score_from_sums(sums, C, HW), x_next
    x_next, u8, _ = stepper.step(x_cur, eps, i, want_x_next=want_x_next, want_u8=True, want_sums=False)
    timesteps = torch.zeros(u8.shape[0], device=u8.device)               # edm/main.py:829
    timesteps._b200_uniform_value = 0.0       # lets the B200 scorers skip the all-equal check (a host sync)
    s = scorer(u8, labels_rows, timesteps)
    return torch.as_tensor(s).to(device=u8.device, dtype=torch.float32).reshape(-1).contiguous(), x_next


def _gather_winner(rows: torch.Tensor, idx: torch.Tensor, lo: int, hi: int, shard: 'Shard') -> torch.Tensor:
    """rows [hi-lo, b, ...] fp64 = this rank's slice; idx [b] global winner indices -> [b, ...] on every rank."""
    if shard.world == 1:
        return ops.gather_rows(rows, (idx - lo).contiguous())
    from ..sharding import exchange_winner
    return exchange_winner(rows, idx, lo, hi, shard.group, gather=ops.gather_rows)


def _scale_table(num_steps: int, K: int, N: int, lam: float) -> torch.Tensor:
    """fp32 scales of edm/main.py:776-779 for every (i,k,n): `ones * scale_seed * lambda_param`
    (fp32, rounded after each product); hash() is this process's salted str hash, as in the reference."""
    seeds = torch.tensor([[[hash(f"{i}_{k}_{n}") % 1000 / 1000.0 for n in range(N)] for k in range(K)]
                          for i in range(num_steps)], dtype=torch.float64)
    return (torch.ones_like(seeds, dtype=torch.float32) * seeds.to(torch.float32)) * torch.tensor(lam).to(torch.float32)


class _NoiseStager:
    """Host -> device staging of precomputed noise ONE round ahead on a side stream, so that the (pinned) host buffers of
    round r+1 cross PCIe underneath the network evaluations of round r.  Two persistent device buffers per tensor kind are
    reused alternately (no allocator traffic, nothing to defer-free across streams): round r+1 is staged into the buffer
    round r-1 used, after an event on the compute stream that follows all of round r-1's work.  Only host-resident fp64
    tensors are staged; everything else falls through to the caller's own path."""

    def __init__(self, pre: Optional[Dict], device, lo: int, hi: int, enabled: bool = True):
        self.pre, self.device, self.lo, self.hi = pre, device, lo, hi
        self.enabled = bool(enabled) and pre is not None and torch.device(device).type == 'cuda'
        self.stream = torch.cuda.Stream(device=device) if self.enabled else None
        self.bufs: Dict[Any, torch.Tensor] = {}
        self.staged: Dict[Any, Any] = {}
        self.round = 0

    def _host(self, key):
        t = self.pre.get(key) if self.pre is not None else None
        return t if (torch.is_tensor(t) and not t.is_cuda and t.dtype == torch.float64) else None

    def prefetch(self, i, k: int):
        if not self.enabled or i is None:
            return
        self.round += 1
        main = torch.cuda.current_stream(self.device)
        fence = torch.cuda.Event()
        fence.record(main)                       # everything enqueued so far (incl. the round that last used the buffer)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(fence)
            for key in ((f'pivot_{i}',) if k == 0 else ()) + ((i, k),):
                if key in self.staged:
                    continue
                src = self._host(key[0] if isinstance(key, tuple) else key)
                if src is None:
                    continue
                if isinstance(key, tuple):
                    if k >= src.shape[1] or self.hi > src.shape[2]:
                        continue
                    src = src[:, k, self.lo:self.hi]
                slot = ('dir' if isinstance(key, tuple) else 'pivot', self.round & 1, tuple(src.shape))
                if slot not in self.bufs:
                    self.bufs[slot] = torch.empty(src.shape, dtype=torch.float64, device=self.device)
                dst = self.bufs[slot]
                dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.staged[key] = (dst, ev)

    def take(self, key):
        """The staged device copy of `key` (made visible to the current stream), or None."""
        item = self.staged.pop(key, None)
        if item is None:
            return None
        dst, ev = item
        torch.cuda.current_stream(self.device).wait_event(ev)
        return dst


# Speculation past escalated rounds (eps_greedy_search `speculate`): B200NS_SPECULATE=1 turns the default on,
# B200NS_SPEC_PRIO is the CUDA priority of the stream the precise pass runs on (default -1 = above the main stream),
# B200NS_SPEC_GAP the minimum lead of the 16-bit winner (in standard deviations of the round's scores) to speculate on.
# OFF by default -- built, bit-identical (tests/test_search_gpu.py::test_speculation_*), and measured on B200 (DESIGN.md 4e,
# profiles/r02_speculation_ab.txt): the 16-bit GEMMs are persistent and hold every SM, so the ~940 dependent launches of a
# precise pass only advance at the main stream's kernel boundaries and the pass takes about as long as the round it hides
# behind: 37.0 vs 37.1 ms per step with three of six escalated rounds speculated (all held), 36.9 vs 37.2 with two; and
# without the lead gate two of six guesses were wrong (leads of 0.008 and 0.03-0.04 std) and cost a round each: 37.7 vs 37.0.
SPECULATE_DEFAULT = os.environ.get('B200NS_SPECULATE', '0') == '1'
SPEC_GAP = float(os.environ.get('B200NS_SPEC_GAP', '0.08'))       # minimum lead of the 16-bit winner, in units of the round's score std
_SPEC_STREAMS: Dict[Any, Any] = {}


def _speculation_on(speculate: Optional[bool], escalate: bool, scorer, device) -> bool:
    if not escalate or torch.device(device).type != 'cuda':
        return False
    want = SPECULATE_DEFAULT if speculate is None else bool(speculate)
    if want and not getattr(scorer, 'fused_sums', False):
        # scorer networks (classifier, CLIP) own static plan buffers: two streams may not score through them at once
        if speculate:
            raise NotImplementedError('speculate=True needs a scorer that works from the fused channel sums (brightness)')
        return False
    return want


def _speculation_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SPEC_STREAMS:
        _SPEC_STREAMS[key] = torch.cuda.Stream(device=device, priority=int(os.environ.get('B200NS_SPEC_PRIO', '-1')))
    return _SPEC_STREAMS[key]


def _key_score(key: torch.Tensor) -> torch.Tensor:
    """fp32 score packed in the high half of an argmax key (csrc/sampler.cuh: orderable_f32 is an involution)."""
    o = (key >> 32).to(torch.int32)
    return torch.where(o < 0, o ^ 0x7FFFFFFF, o).view(torch.float32)


def _escalate(scorer, stepper: HeunStepper, x_cur, local, scores, key, idx, i, lo, hi, b, labels_rows, C, HW, shard,
              delta: Optional[float], kappa: float, max_rows: int, side=None, spec_gap: float = 0.0,
              log: Optional[list] = None):
    """One round's near-tie escalation.  scores [nl, b] (this rank's bf16 scores), key [b] (the GLOBAL packed argmax key).
    Returns (idx [b] global winners, rows refined here, refined score table or None, truncated?).

    Contenders of image j: every candidate whose score is within delta_j of the round's best, delta_j = `delta` if given,
    else kappa * std_n(scores[:, j]) (the bf16 score noise that matters -- the part that DIFFERS between candidates --
    measures ~0.07 of the spread of the scores themselves at every noise level), but never more than the
    `max_rows` best-scoring ones per image.  All of this is a function of the GLOBAL score table, so a sharded run refines
    exactly the rows an unsharded run refines.

    `side` (a CUDA stream): SPECULATIVE mode, five values are returned.  If the 16-bit winner leads the runner-up by at least
    `spec_gap` x the standard deviation of the round's scores (a narrower lead is too likely to be overturned -- the
    difference of two candidates' 16-bit errors has sigma ~ 0.1 std -- and a wrong guess costs a whole round), the precise
    pass and the argmax over the refined table are enqueued on `side`, the fifth value is a `_Pending` and `idx` comes back
    unchanged (the 16-bit winner, provisional): the caller carries on with it while the precise pass runs underneath the
    next round, and `_Pending.resolve()` later tells whether the refined argmax agreed.  Otherwise the pass runs
    synchronously as without `side` and the fifth value is None."""
    nl = hi - lo
    N = nl * shard.world
    best = _key_score(key)                                                  # [b] global best score
    dist = None
    if shard.world > 1:
        import torch.distributed as dist
    std = None
    if delta is None or side is not None:
        mom = torch.stack([scores.sum(0, dtype=torch.float64), (scores.double() ** 2).sum(0)])        # [2, b]
        if dist is not None:
            dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=shard.group)
        mean = mom[0] / N
        std = (mom[1] / N - mean * mean).clamp_min(0).sqrt().to(torch.float32)
    d = kappa * std if delta is None else torch.full_like(best, float(delta))
    thr = best - d
    # the max_rows-th best score per image, globally: the contender list never grows beyond max_rows per image
    M = min(max_rows, N)
    top = torch.topk(scores, min(M, nl), dim=0).values                      # [<=M, b] local
    if dist is not None:
        allt = [torch.empty_like(top) for _ in range(shard.world)]
        dist.all_gather(allt, top.contiguous(), group=shard.group)
        top = torch.topk(torch.cat(allt, dim=0), M, dim=0).values
    floor = top[M - 1]
    mask = scores >= torch.maximum(thr, floor).unsqueeze(0)                 # contenders (the best itself included)
    cnt = mask.sum(dim=0)
    wide = (scores >= thr.unsqueeze(0)).sum(dim=0)                          # without the cap
    if dist is not None:
        both = torch.stack([cnt, wide])
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=shard.group)
        cnt, wide = both[0], both[1]
    multi = cnt > 1                                                         # images whose best has company
    sel = mask & multi.unsqueeze(0)
    # speculation gate: every image whose best has company must be led by >= spec_gap x std (top = the global top-M scores)
    clear = ((top[0] - top[1] >= spec_gap * std) | ~multi) if (side is not None and M > 1) else torch.zeros_like(multi)
    host = torch.cat([sel.reshape(-1).to(torch.float32), (wide > cnt).to(torch.float32), multi.to(torch.float32),
                      clear.to(torch.float32), ((top[0] - top[1]) / std.clamp_min(1e-30)) if (std is not None and M > 1)
                      else torch.zeros_like(best)]).cpu()                                              # THE host sync
    m_host = host[:nl * b].bool()
    truncated = bool(host[nl * b:nl * b + b].any())
    if not bool(host[nl * b + b:nl * b + 2 * b].any()):
        return (idx, 0, None, False) if side is None else (idx, 0, None, False, None)
    speculative = side is not None and bool(host[nl * b + 2 * b:nl * b + 3 * b].all())
    if log is not None:       # (timestep, lead of the 16-bit winner over the runner-up in score standard deviations, speculated?)
        log.append((i, [round(float(v), 4) for v in host[nl * b + 3 * b:]], speculative))
    rows = torch.nonzero(m_host).flatten()                                  # row = n_local * b + j
    refined = torch.where(mask & ~multi.unsqueeze(0), scores, torch.full_like(scores, float('-inf')))
    n_rows = int(rows.numel())

    def precise_pass():
        if n_rows:
            rows_d = rows.to(scores.device)
            eps_rows = local.index_select(0, rows_d).contiguous()
            lab = labels_rows.index_select(0, rows_d) if labels_rows is not None else None
            row_images = (rows_d % b) if b > 1 else None
            if getattr(scorer, 'fused_sums', False):
                _, _, sums = stepper.step(x_cur, eps_rows, i, want_x_next=False, want_sums=True, precise=True, row_images=row_images)
                s_p = scorer.score_from_sums(sums, C, HW)
            else:
                _, u8, _ = stepper.step(x_cur, eps_rows, i, want_x_next=False, want_u8=True, want_sums=False, precise=True,
                                        row_images=row_images)
                timesteps = torch.zeros(u8.shape[0], device=u8.device)
                timesteps._b200_uniform_value = 0.0
                s_p = torch.as_tensor(scorer(u8, lab, timesteps)).to(device=u8.device, dtype=torch.float32).reshape(-1)
            refined.view(-1).index_copy_(0, rows_d, s_p.contiguous())
        return ops.argmax_first(refined.contiguous(), idx_base=lo, want_key=True)

    def reduce_key(idx2, key2):
        if dist is not None:
            dist.all_reduce(key2, op=dist.ReduceOp.MAX, group=shard.group)
            return 0xFFFFFFFF - (key2 & 0xFFFFFFFF)
        return idx2 + lo

    if not speculative:
        out = (reduce_key(*precise_pass()), n_rows, refined, truncated)
        return out if side is None else out + (None,)
    main = torch.cuda.current_stream(scores.device)
    ready = torch.cuda.Event()
    ready.record(main)                     # everything the precise pass reads (x_cur, local, refined) has been enqueued
    with torch.cuda.stream(side):
        side.wait_event(ready)
        idx2, key2 = precise_pass()
        done = torch.cuda.Event()
        done.record(side)
    return idx, n_rows, refined, truncated, _Pending(done, idx2, key2, idx, reduce_key, keep=(x_cur, local, refined, scores))


class _Pending:
    """A precise re-scoring pass in flight on the speculation stream (see eps_greedy_search `speculate`)."""

    def __init__(self, done, idx2, key2, idx_prov, reduce_key, keep):
        self.done, self.idx2, self.key2, self.idx_prov, self.reduce_key, self.keep = done, idx2, key2, idx_prov, reduce_key, keep
        self.ctx = None                    # the round's finish context, filled in by the search loop

    def resolve(self) -> Optional[torch.Tensor]:
        """Wait for the precise pass; None if the refined argmax equals the provisional (16-bit) winner, else the
        refined winner indices [b] (global).  One host synchronisation."""
        torch.cuda.current_stream(self.idx2.device).wait_event(self.done)
        idx_true = self.reduce_key(self.idx2, self.key2)
        return None if torch.equal(idx_true, self.idx_prov) else idx_true


@torch.no_grad()
def eps_greedy_search(net: B200Denoiser, latents, class_labels, params: SamplingParams, table: StepTable, *,
                      precomputed_noise: Optional[Dict] = None, shard: Optional[Shard] = None, record: bool = False,
                      norm_mode: str = 'kernel', scale_table: Optional[torch.Tensor] = None,
                      teacher_x: Optional[List[torch.Tensor]] = None, step_indices: Optional[List[int]] = None,
                      x_init: Optional[torch.Tensor] = None, on_step=None,
                      commit: str = 'reuse', mirror_rng: bool = True, prefetch: bool = False,
                      dedupe_noise_free: bool = False, escalate: Optional[bool] = None, delta: Optional[float] = None,
                      kappa: float = ESCALATION_KAPPA, max_contenders: int = MAX_CONTENDERS,
                      bernoulli_draws: Optional[torch.Tensor] = None, speculate: Optional[bool] = None,
                      spec_gap: Optional[float] = None, _spec_sabotage: bool = False) -> (torch.Tensor, SearchRecord):
    """ZERO_ORDER == EPS_GREEDY branch (edm/main.py:714-860).

    Extras over the reference (all optional): `shard` (candidate sharding over ranks), `record`,
    `norm_mode` ('kernel' = in-kernel fp64 block reduction, 'torch' = the reference's own torch.norm
    calls), `teacher_x` (force the committed state per step; parity tests), `step_indices` / `x_init`
    (run a sub-sequence of steps from a given state; benchmarks), `on_step(i, x_next, idx, scores)`
    (called after every committed step, e.g. to read results back to the host).

    `dedupe_noise_free` (off by default; never used for the headline number): on the steps whose noise scale is exactly 0
    (t outside [S_min, S_max]) all N candidates are the same tensor -- evaluate one and replicate its score / state.
    Bit-identical results; bench.py reports the throughput with it as a separate figure.

    `prefetch`: precomputed noise that lives in (pinned) host memory is staged one round ahead on a side stream, so the
    host->device copies run underneath the previous round's network evaluations.  Meant for callers whose `on_step` does
    not synchronise (asynchronous copies into pinned buffers, as bench.py's end-to-end leg does): measured on B200,
    33.27 -> 32.96 ms per step end to end; with a blocking `on_step` (`.cpu()`) it is counter-productive (42 ms), hence off
    by default.

    `escalate` (default: on whenever the network has a precise twin, i.e. ADM): near-tie precision escalation.  The bf16
    network puts ~1e-4 of noise on the scores while the reference runs it in fp32, so after the bf16 pass every candidate
    whose score is within delta of the round's best -- `delta` if given, else `kappa` x the standard deviation of the
    round's scores (the candidate-dependent part of the bf16 score noise is a fixed fraction of that spread) -- is
    re-evaluated by the split-fp16 (fp32-faithful) engine and the first-max argmax is taken over the refined scores
    (edm/main.py:842 on the reference's own precision).  Rounds with a single contender -- and the noise-free steps, whose
    N candidates are one tensor -- cost nothing extra; otherwise one host read per round decides which rows to refine (at
    most `max_contenders` per image, best first).  The committed state stays the bf16 engine's x_next of the winner, so
    commit 'reuse' and 'recompute' remain bit-identical.

    `speculate` (default: B200NS_SPECULATE, off -- measured to gain < 1 %, see SPECULATE_DEFAULT): with escalation on and a scorer that works from the fused channel sums
    (brightness), an escalated round does not wait for its precise pass.  The pass runs on a second, high-priority stream
    while the main stream starts the next round from the provisional 16-bit winner; after that round's scoring the refined
    argmax is compared with the provisional one (the one host read an escalated round costs anyway).  Agreement -- every
    round of the reference fixtures with IEEE-half storage -- means the ~14 ms latency-bound precise pass was hidden behind
    throughput-bound work; disagreement rolls back: the escalated round's pivot / commit / trace are redone with the
    refined winner and the next round is re-run on the noise inputs already prepared for it (they do not depend on any winner,
    so no RNG state has to be restored).  Either way the results are bit-identical to
    the synchronous path (tests/test_search_gpu.py::test_speculation_*).  `on_step` of a step that rests on an unverified
    winner is delivered once it is verified (one round later).  `spec_gap` (default B200NS_SPEC_GAP = 0.08): an escalated
    round is only speculated past when its 16-bit winner leads the runner-up by that many standard deviations of the round's
    scores; narrower leads are overturned too often (a wrong guess costs a round) and wait for their precise pass.

    `bernoulli_draws` (tests): the uniform draws of edm/main.py:751 in call order, [num_steps*K*N] fp32, used INSTEAD of
    `torch.rand(1, device)` -- a reference run on another device (CPU) drew them from another generator.

    `commit`: the reference re-runs `step` on the winning noise at batch b (edm/main.py:860).  All kernels
    here are batch-size and batch-position invariant (fixed reduction orders), so the winner's x_next from
    the last candidate round is bit-identical to that recomputation ('reuse', default; asserted by
    tests/test_search_gpu.py::test_commit_reuse_is_bit_identical); 'recompute' runs the two extra network
    evaluations like the reference."""
    if commit not in ('reuse', 'recompute'):
        raise ValueError("commit must be 'reuse' or 'recompute'")
    device = net.device
    shard = shard or Shard()
    N, K, eps_p = params.N, params.K, params.eps
    lo, hi = shard.bounds(N)
    lam = params.lambda_param * np.sqrt(3 * 64 * 64)                      # :716
    num_steps = table.num_steps
    rec = SearchRecord()
    if x_init is not None:
        x_next = x_init.to(device=device, dtype=torch.float64).contiguous()
    else:
        x_next = latents.to(torch.float64) * table.t_steps[0]             # :99
    b = x_next.shape[0]
    C, HW = x_next.shape[1], x_next.shape[2] * x_next.shape[3]
    stepper = HeunStepper(net, table, class_labels)
    scales = (scale_table if scale_table is not None else _scale_table(num_steps, K, N, lam)).to(device)
    if shard.world > 1:
        # hash() is salted per process (edm/main.py:776 inherits that): all ranks must perturb with rank 0's scales
        import torch.distributed as dist
        scales = scales.contiguous()
        dist.broadcast(scales, src=dist.get_global_rank(shard.group, 0) if shard.group is not None else 0, group=shard.group)
    labels_rows = class_labels.repeat(hi - lo, 1) if class_labels is not None else None
    use_mirror = mirror_rng and philox.mirror_ok(device)
    if escalate is None:
        escalate = net.supports_precise
    elif escalate and not net.supports_precise:
        raise NotImplementedError('escalate=True: the precise engine implements head_dim-64 attention (ADM) only')
    if escalate:
        for R in range(1, min(max_contenders, MAX_CONTENDERS) + 1):   # the usual contender batch sizes up front: no plan is built mid-run
            net.precise_engine.plan(R, 1 if b == 1 else R)
    pre = precomputed_noise
    if pre is not None and 'pivot' in pre:                                # :724-727 (value unused, RNG untouched)
        pass
    else:
        torch.randn_like(x_next)                                          # keeps the RNG stream aligned with :727
    seq = list(step_indices) if step_indices is not None else list(range(num_steps))
    stager = _NoiseStager(pre, device, lo, hi, enabled=prefetch)
    # ---- speculation past an escalated round (`speculate`): the precise re-scoring of round r runs on a second stream
    # underneath the 16-bit evaluations of round r+1, which start from the provisional (16-bit) winner; the refined argmax
    # is checked right after round r+1's scoring and, if it disagrees, round r's outcome is corrected and round r+1 re-run
    # on its (winner-independent, already prepared) noise inputs -- results are identical to the synchronous path either way.
    spec = _speculation_on(speculate, escalate, params.scorer, device)
    side = _speculation_stream(device) if spec else None
    rounds = [(pos, i, k) for pos, i in enumerate(seq) for k in (range(K) if K > 0 else (None,))]
    pending: Optional[_Pending] = None
    held: List[tuple] = []               # on_step calls of rounds that still rest on an unverified winner
    x_cur = pivot = None

    def finish(ctx, idx, hold: bool):
        """Everything of a round that depends on its winner `idx`: the new pivot, the commit at the end of a step, the
        trace.  Returns (pivot, x_next); called again with the refined winner when a speculation failed."""
        i, k, x_cur, local, x_cands, scores, pivot = (ctx[n] for n in ('i', 'k', 'x_cur', 'local', 'x_cands', 'scores', 'pivot'))
        x_next = ctx['x_next']
        if k is not None:
            rec.escalated.append(ctx['n_esc'])
            if record:
                rec.refined.append(ctx['refined'])
            # ---- new pivot = the winning candidate (:848-857); it lives on its owner's GPU only: the other ranks
            # contribute zeros to an all_reduce(SUM) (x + 0 is exact), 96 KiB per image over NVLink -- and only when
            # somebody reads it (another local-search round, a recomputed commit, a trace)
            if k < K - 1 or commit != 'reuse' or record:
                pivot = _gather_winner(local.reshape(hi - lo, b, *local.shape[1:]), idx, lo, hi, shard)
            if record:
                rec.scores.append(scores)
                rec.indices.append(idx)
            if k < K - 1:
                return pivot, x_next
        # ---- commit (:860)
        if commit == 'reuse' and K > 0:
            x_next = _gather_winner(x_cands.reshape(hi - lo, b, *x_cands.shape[1:]), idx, lo, hi, shard)
        else:
            x_next, _, _ = stepper.step(x_cur, pivot, i, want_x_next=True)
        if record:
            rec.pivots.append(pivot)
            rec.x_steps.append(x_next)
        if on_step is not None:
            (held.append if hold else (lambda a: on_step(*a)))((i, x_next, idx, scores))
        if teacher_x is not None:
            x_next = teacher_x[i].to(device=device, dtype=torch.float64).contiguous()
        return pivot, x_next

    def rec_mark():
        return tuple(len(l) for l in (rec.scores, rec.indices, rec.escalated, rec.refined, rec.pivots, rec.x_steps))

    def rec_rewind(mark):
        for l, n in zip((rec.scores, rec.indices, rec.escalated, rec.refined, rec.pivots, rec.x_steps), mark):
            del l[n:]

    def settle(p: _Pending) -> bool:
        """Verify the speculation `p`; on a miss redo its round's `finish` with the refined winner.  True = it held."""
        nonlocal pivot, x_next
        idx_true = p.resolve()
        if idx_true is None:
            for call in held:
                on_step(*call)
            held.clear()
            return True
        rec.mispredicted += 1
        rec.missed_steps.append(p.ctx['i'])
        held.clear()
        rec_rewind(p.ctx['mark'])
        pivot, x_next = finish(p.ctx, idx_true, False)
        return False

    # ---- a round's noise inputs (pivot draw of a new step, directions, fresh noises, Bernoulli mask, direction norms) do not
    # depend on any winner: they are prepared -- RNG draws in the reference's order, uploads, norms -- one round AHEAD, right
    # after the previous round's evaluation has been enqueued and before its escalation decision synchronises with the host,
    # so that the host-side work (64 dictionary lookups and a stack per round with eps > 0, the Philox mirror, ...) runs
    # underneath the GPU's work instead of in the bubble behind the synchronisation (0.5-0.7 ms of host time per round).
    # A re-run round (speculation miss) reuses its prepared inputs: no RNG state to restore.
    tmpl = x_next                        # shape / dtype of one image batch: fp64 [b, C, H, W]
    prepared: Dict[int, dict] = {}

    def prepare(r: int) -> dict:
        pos, i, k = rounds[r]
        out = {}
        x_cur = tmpl                     # only shapes / dtypes are taken from these two below
        if k is None or k == 0:
            if pre is not None and f'pivot_{i}' in pre:                       # :734-737
                pivot = stager.take(f'pivot_{i}')
                if pivot is not None:
                    pivot = pivot.clone()             # the staging buffer is recycled two rounds later; pivots may be recorded
                else:
                    pivot = pre[f'pivot_{i}'].to(device=device, dtype=torch.float64, non_blocking=True)
                pivot = pivot.contiguous()
            else:
                pivot = torch.randn_like(x_cur)
            out['pivot0'] = pivot
        if k is None:
            return out
        pivot = tmpl
        # next round's host noise starts crossing PCIe now, on the side stream
        stager.prefetch(*((i, k + 1) if k + 1 < K else (seq[pos + 1] if pos + 1 < len(seq) else None, 0)))
        # ---- candidate construction (:749-800).  RNG calls mirror the reference one for one; the
        # Bernoulli stays on the device (no host sync) unless precomputed noise covers only one of
        # the two branches AND 0 < eps < 1, where the reference's RNG consumption is data dependent.
        bulk = (pre is not None and i in pre and k < pre[i].shape[1] and N <= pre[i].shape[2] and
                (eps_p <= 0 or all(f'fresh_{i}_{k}_{n}' in pre for n in range(N))))
        dirs, fresh, perturb = [], [], []
        if bulk and bernoulli_draws is not None:
            r0 = (i * K + k) * N
            perturb_host = bernoulli_draws[r0:r0 + N].to(torch.float32).cpu().numpy() < np.float32(1 - eps_p)
        elif bulk and use_mirror:
            # the N `torch.rand(1)` draws of :751 evaluated on the host from the generator's (seed, offset) --
            # same values, same final RNG state, no kernels in the stream (philox.py)
            perturb_host = philox.rand1_sequence(device, N) < np.float32(1 - eps_p)
        else:
            perturb_host = None
            for n in range(N):
                p_t = torch.rand(1, device=device) < (1 - eps_p)                  # :751
                perturb.append(p_t)
                if bulk:
                    continue
                has_dir = pre is not None and i in pre and k < pre[i].shape[1] and n < pre[i].shape[2]
                fkey = f'fresh_{i}_{k}_{n}'
                has_fresh = pre is not None and fkey in pre
                z_dir = pre[i][:, k, n].reshape(pivot.shape) if has_dir else None          # :755-759
                z_fresh = pre[fkey] if has_fresh else None                                 # :791-792
                if not has_dir and not has_fresh:
                    z_dir = z_fresh = torch.randn_like(pivot)             # :767 / :795: one draw either way
                elif not (has_dir and has_fresh):
                    branch = True if eps_p <= 0 else (False if eps_p >= 1 else bool(p_t))
                    if branch and not has_dir:
                        z_dir = torch.randn_like(pivot)
                    if not branch and not has_fresh:
                        z_fresh = torch.randn_like(x_cur)
                    z_dir = z_dir if z_dir is not None else z_fresh
                    z_fresh = z_fresh if z_fresh is not None else z_dir
                dirs.append(z_dir)
                fresh.append(z_fresh)
        # ---- only this rank's candidates [lo, hi) are materialised (1/G of the transfers and of the fp64 passes)
        nl = hi - lo
        if bulk:      # every direction comes from one precomputed tensor: a single (async) transfer of the slice
            Z = stager.take((i, k))
            if Z is None:
                Z = pre[i][:, k, lo:hi].to(device=device, dtype=torch.float64, non_blocking=True)
            Z = Z.transpose(0, 1).reshape(nl * b, *pivot.shape[1:]).contiguous()
            ZF = Z if eps_p <= 0 else torch.stack([pre[f'fresh_{i}_{k}_{n}'].to(device=device, dtype=torch.float64)
                                                   for n in range(lo, hi)]).reshape(nl * b, *pivot.shape[1:]).contiguous()
        else:
            as64 = lambda ts: torch.stack([t.to(device=device, dtype=torch.float64) for t in ts]).reshape(
                nl * b, *pivot.shape[1:]).contiguous()
            Z = as64(dirs[lo:hi])
            ZF = Z if all(f is d for f, d in zip(fresh[lo:hi], dirs[lo:hi])) else as64(fresh[lo:hi])
        if perturb_host is not None:
            if perturb_host[lo:hi].all():
                fresh_mask = torch.zeros(nl * b, dtype=torch.uint8, device=device)
            else:
                # pinned + non_blocking: a copy from pageable memory is a synchronous cudaMemcpy (the pinned block is recycled by
                # torch's caching host allocator only after the copy has completed)
                fresh_mask = torch.from_numpy((~perturb_host[lo:hi]).astype(np.uint8)).pin_memory().to(
                    device, non_blocking=True).repeat_interleave(b).contiguous()
        else:
            fresh_mask = (~torch.cat(perturb[lo:hi])).to(torch.uint8).repeat_interleave(b).contiguous()
        if norm_mode == 'torch':                                      # strict: the reference's own call (:764)
            norms = torch.cat([torch.norm(z, p=2, dim=tuple(range(1, z.dim()))) for z in Z.reshape(nl, b, *pivot.shape[1:])]).to(torch.float64)
        else:
            norms = ops.direction_norms(Z)
        sc = scales[i, k, lo:hi].repeat_interleave(b).contiguous()
        out.update(Z=Z, ZF=ZF, fresh_mask=fresh_mask, norms=norms, sc=sc)
        return out

    r = 0
    while r < len(rounds):
        pos, i, k = rounds[r]
        if r not in prepared:
            prepared[r] = prepare(r)
        inp = prepared[r]
        if k is None or k == 0:
            x_cur = x_next
            pivot = inp['pivot0']
        if k is None:                         # K == 0: no search rounds, the pivot is committed as drawn
            pivot, x_next = finish(dict(i=i, k=None, x_cur=x_cur, local=None, x_cands=None, scores=None, pivot=pivot,
                                        x_next=x_next), None, False)
            prepared.pop(r, None)
            r += 1
            continue
        nl = hi - lo
        local = ops.make_candidates(pivot, inp['Z'], inp['norms'], inp['sc'], inp['fresh_mask'], inp['ZF'])      # [(hi-lo)*b, C, H, W]
        # ---- evaluate this rank's slice: 2 NFE + Tweedie x0 + score (:809-838)
        want_x = commit == 'reuse' and k == K - 1
        if dedupe_noise_free and table.steps[i].s == 0.0:
            # gamma = 0: x_hat = x_cur for every candidate (edm/main.py:83-85), so the N candidates are the same
            # tensor; the kernels are batch-position invariant, hence N identical scores and identical x_next
            # (tests: test_noise_free_steps_are_exact_ties).  Evaluate candidate 0 and replicate -- bit-identical.
            s1, x1 = _score_rows(params.scorer, stepper, x_cur, local[:b], i,
                                 labels_rows[:b] if labels_rows is not None else None, C, HW, want_x)
            scores = s1.reshape(1, b).expand(hi - lo, b).contiguous()
            x_cands = x1.unsqueeze(0).expand(hi - lo, *x1.shape).reshape((hi - lo) * b, *x1.shape[1:]).contiguous() if want_x else None
        else:
            scores, x_cands = _score_rows(params.scorer, stepper, x_cur, local, i, labels_rows, C, HW, want_x)
            scores = scores.reshape(hi - lo, b)
        rec.scored_candidates += (hi - lo) * b
        # ---- first-max argmax (+ cross-rank reduction of the packed key) (:842)
        idx, key = ops.argmax_first(scores, idx_base=lo, want_key=True)
        if shard.world > 1:
            import torch.distributed as dist
            dist.all_reduce(key, op=dist.ReduceOp.MAX, group=shard.group)
            idx = 0xFFFFFFFF - (key & 0xFFFFFFFF)
        else:
            idx = idx + lo
        # ---- the previous round's speculation is settled here: its precise pass ran underneath this round's evaluations
        if pending is not None:
            p, pending = pending, None
            if not settle(p):
                # the refined winner differs: pivot / x_next have been corrected, this round ran on the wrong state -> redo
                rec.scored_candidates -= (hi - lo) * b
                continue
        if r + 1 < len(rounds) and (r + 1) not in prepared:
            prepared[r + 1] = prepare(r + 1)          # the next round's noise inputs, ahead of the host synchronisation below
        # ---- near-tie escalation: re-score the contenders with the fp32-faithful engine, argmax over the refined table
        n_esc, refined = 0, None
        if escalate and table.steps[i].s != 0.0:
            if spec:
                idx, n_esc, refined, trunc, pending = _escalate(params.scorer, stepper, x_cur, local, scores, key, idx, i, lo, hi,
                                                                b, labels_rows, C, HW, shard, delta, kappa, max_contenders,
                                                                side=side, spec_gap=SPEC_GAP if spec_gap is None else spec_gap,
                                                                log=rec.escalation_log)
                if pending is not None and _spec_sabotage:
                    idx = (idx + 1) % N          # tests: a wrong provisional winner forces the rollback path
                    pending.idx_prov = idx
            else:
                idx, n_esc, refined, trunc = _escalate(params.scorer, stepper, x_cur, local, scores, key, idx, i, lo, hi, b,
                                                       labels_rows, C, HW, shard, delta, kappa, max_contenders)
            rec.truncated += int(trunc)
        ctx = dict(i=i, k=k, x_cur=x_cur, local=local, x_cands=x_cands, scores=scores, pivot=pivot, x_next=x_next,
                   n_esc=n_esc, refined=refined, mark=rec_mark())
        if pending is not None:
            pending.ctx = ctx
        pivot, x_next = finish(ctx, idx, pending is not None)
        prepared.pop(r, None)
        r += 1
    if pending is not None:
        settle(pending)
    return x_next, rec


@torch.no_grad()
def naive_search(net, latents, class_labels, table: StepTable, *, noise: Optional[List[torch.Tensor]] = None,
                 record=False):
    """NAIVE branch (edm/main.py:862-866)."""
    rec = SearchRecord()
    stepper = HeunStepper(net, table, class_labels)
    x_next = latents.to(torch.float64) * table.t_steps[0]
    for i in range(table.num_steps):
        eps_i = noise[i].to(device=net.device, dtype=torch.float64).contiguous() if noise is not None else torch.randn_like(x_next)
        x_next, _, _ = stepper.step(x_next, eps_i, i, want_x_next=True)
        if record:
            rec.x_steps.append(x_next)
    return x_next, rec


@torch.no_grad()
def rejection_search(net, latents, class_labels, params: SamplingParams, table: StepTable, *,
                     precomputed_noise: Optional[Dict] = None, record=False):
    """REJECTION_SAMPLING branch (edm/main.py:101-137): rows are image-major (row = j*N + n)."""
    N = params.N
    rec = SearchRecord()
    x_next = latents.to(torch.float64) * table.t_steps[0]
    b = x_next.shape[0]
    x = x_next.repeat_interleave(N, dim=0).contiguous()
    labels = class_labels.repeat_interleave(N, dim=0) if class_labels is not None else None
    stepper = HeunStepper(net, table, labels)
    for i in range(table.num_steps):
        if precomputed_noise is not None and i in precomputed_noise:
            eps_i = precomputed_noise[i][:, :N].reshape(b * N, *x.shape[1:]).to(device=net.device, dtype=torch.float64).contiguous()
        else:
            eps_i = torch.randn_like(x)
        x, _, _ = stepper.step(x, eps_i, i, want_x_next=True)
    u8 = ops.quantize_u8(x)
    scores = params.scorer(u8, labels, torch.zeros(u8.shape[0], device=u8.device))
    scores = torch.as_tensor(scores).to(device=net.device, dtype=torch.float32).view(b, N)
    rec.scored_candidates = b * N
    best = ops.argmax_first(scores.t().contiguous())                      # [N,b] -> [b]
    xr = x.view(b, N, *x.shape[1:]).transpose(0, 1).contiguous()          # [N,b,...]
    x_next = ops.gather_rows(xr, best)
    if record:
        rec.scores.append(scores)
        rec.indices.append(best)
    return x_next, rec


class _MCTSNode:
    __slots__ = ('x', 'depth', 'children', 'reward', 'visit')

    def __init__(self, x, depth, visit=0):
        self.x, self.depth, self.children, self.reward, self.visit = x, depth, [], 0.0, visit


def _mcts_select(root: _MCTSNode) -> List[_MCTSNode]:
    """UCB1 descent (edm/main.py:548-572): unvisited children score +inf, np.argmax takes the first maximum."""
    path, node = [root], root
    while node.children:
        ucb = [float('inf') if ch.visit == 0 else ch.reward / ch.visit + np.sqrt(2 * np.log(node.visit) / ch.visit)
               for ch in node.children]
        node = node.children[int(np.argmax(ucb))]
        path.append(node)
    return path


@torch.no_grad()
def mcts_search(net: B200Denoiser, latents, class_labels, params: SamplingParams, table: StepTable, *,
                precomputed_noise: Optional[Dict] = None, record: bool = False):
    """MCTS branch (edm/main.py:405-713): b = params.N children per expansion, params.S simulations per
    timestep in groups of 16 whose rewards are backed up after the whole group, deterministic (zero-noise) rollouts to
    t = 0 scored on the final image, root <- visited child with the best mean reward (first maximum), subtree kept.
    The tree logic and both RNG streams (torch.randn for the per-depth noises and the throw-away draw of every expansion,
    np.random.randint for the rollout child) follow the reference call for call; what changes is the execution:
      * an expansion is ONE Heun step at batch b (the reference steps the b children of a deeper node one by one),
      * the rollouts of a simulation group advance together -- at step j one batched Heun step over every simulation that
        has reached depth j (the reference finishes them one by one at batch 1); rows are padded to 1/2/4/8/16 so that five
        cached plans serve every group.  The engine is batch-size and batch-position invariant, so neither changes a bit.
    Samples are processed in the reference's mini-batches of min(2, batch) (a group's 16 simulations are dealt round-robin
    to the samples of a mini-batch)."""
    device = net.device
    num_steps = table.num_steps
    scorer = params.scorer
    b, S = params.N, params.S
    x_all = latents.to(torch.float64) * table.t_steps[0]
    batch = x_all.shape[0]
    shape = tuple(x_all.shape[1:])
    rec = SearchRecord()
    rec.mcts_rewards, rec.mcts_depths, rec.mcts_chosen = [], [], []
    results = []
    mbs = min(2, batch)
    for mb0 in range(0, batch, mbs):
        mb = min(mb0 + mbs, batch) - mb0
        steppers = [HeunStepper(net, table, None if class_labels is None else class_labels[mb0 + s:mb0 + s + 1])
                    for s in range(mb)]
        depth_noise = {}
        for i in range(num_steps):                                            # :439-447
            if precomputed_noise is not None and i in precomputed_noise:
                depth_noise[i] = precomputed_noise[i].to(device).repeat(mb, 1, 1, 1, 1)
            else:
                depth_noise[i] = torch.randn(mb, b, *shape, device=device)
        roots = [_MCTSNode(x_all[mb0 + s:mb0 + s + 1].clone(), 0, visit=1) for s in range(mb)]

        def expand(node: _MCTSNode, s: int, throwaway: bool):
            i = node.depth
            if throwaway:
                for _ in range(b):
                    torch.randn(1, *shape, device=device)                     # eager default of the dict .get at :578
            eps = depth_noise[i][s].contiguous()     # [b, C, H, W], fp32 like the reference's: the noise term is an fp32 product
            xc, _, _ = steppers[s].step(node.x.contiguous(), eps, i, want_x_next=True)
            node.children = [_MCTSNode(xc[n:n + 1], i + 1) for n in range(b)]

        for i in range(num_steps):
            for s in range(mb):                                               # :467-512
                if not roots[s].children:
                    expand(roots[s], s, throwaway=False)
            group = min(16, S * mb)
            for g0 in range(0, S * mb, group):
                paths, starts = [], []
                for sim in range(g0, min(g0 + group, S * mb)):
                    s = sim % mb
                    path = _mcts_select(roots[s])
                    leaf = path[-1]
                    if leaf.depth < num_steps - 1:                            # :575
                        expand(leaf, s, throwaway=True)
                        leaf = leaf.children[np.random.randint(0, len(leaf.children))]
                        path.append(leaf)
                    paths.append(path)
                    starts.append((leaf.x, leaf.depth, s))
                # ---- deterministic rollouts (:617-640), batched per sample and per step
                finals: List[Optional[torch.Tensor]] = [None] * len(starts)
                for s in range(mb):
                    mine = [k for k, st in enumerate(starts) if st[2] == s]
                    if not mine:
                        continue
                    cur = {k: starts[k][0] for k in mine}
                    for j in range(min(starts[k][1] for k in mine), num_steps):
                        act = [k for k in mine if starts[k][1] <= j]
                        R = len(act)
                        Rp = 1 << (R - 1).bit_length()                        # 1, 2, 4, 8, 16 rows
                        X = torch.cat([cur[k] for k in act] + [cur[act[-1]]] * (Rp - R)).contiguous()
                        xn, _, _ = steppers[s].step(X, torch.zeros_like(X), j, want_x_next=True)
                        for r, k in enumerate(act):
                            cur[k] = xn[r:r + 1]
                    for k in mine:
                        finals[k] = cur[k]
                img = ops.quantize_u8(torch.cat(finals).contiguous())         # :657
                if getattr(scorer, 'fused_sums', False):
                    rewards = scorer.score_from_sums(ops.channel_sums_u8(img), img.shape[1], img.shape[2] * img.shape[3])
                else:
                    labs = None if class_labels is None else torch.cat([class_labels[mb0 + st[2]:mb0 + st[2] + 1] for st in starts])
                    timesteps = torch.zeros(img.shape[0], device=device)
                    timesteps._b200_uniform_value = 0.0
                    rewards = torch.as_tensor(scorer(img, labs, timesteps)).to(device)
                rewards_host = rewards.float().cpu()                          # the tree statistics live on the host (:676)
                rec.scored_candidates += len(starts)
                if record:
                    rec.mcts_rewards.append(rewards_host)
                    rec.mcts_depths.append([st[1] for st in starts])
                for path, r in zip(paths, rewards_host):                      # backup after the whole group (:664-679)
                    for node in path:
                        node.reward += r.item()
                        node.visit += 1
            for s in range(mb):                                               # :682-700
                best, best_r = None, -float('inf')
                for ch in roots[s].children:
                    if ch.visit > 0 and ch.reward / ch.visit > best_r:
                        best, best_r = ch, ch.reward / ch.visit
                assert best is not None
                if record:
                    rec.mcts_chosen.append(roots[s].children.index(best))
                    rec.x_steps.append(best.x)
                roots[s] = best
        results.extend(r.x for r in roots)
    return torch.cat(results), rec


@torch.no_grad()
def beam_search(net, latents, class_labels, params: SamplingParams, table: StepTable, *,
                precomputed_noise: Optional[Dict] = None, record=False):
    """BEAM_SEARCH, intended semantics of edm/main.py:138-404 (k = B beams, N noises per beam,
    score the step's denoised x0, keep the top-k of k*N by score; ties -> lowest flat index
    beam*N + n, as the SD twin's stable sort, pipeline_stable_diffusion.py:1132)."""
    kb, N = params.B, params.N
    rec = SearchRecord()
    x0 = latents.to(torch.float64) * table.t_steps[0]
    b = x0.shape[0]
    C, HW = x0.shape[1], x0.shape[2] * x0.shape[3]
    beams = x0.unsqueeze(0).repeat(kb, 1, 1, 1, 1).reshape(kb * b, *x0.shape[1:]).contiguous()     # [k*b]
    stepper = HeunStepper(net, table, class_labels.repeat(kb, 1) if class_labels is not None else None)
    labels_rows = class_labels.repeat(N * kb, 1) if class_labels is not None else None
    for i in range(table.num_steps):
        if precomputed_noise is not None and i in precomputed_noise:      # [b, k, N, C, H, W]
            eps_i = precomputed_noise[i][:, :kb, :N].permute(2, 1, 0, 3, 4, 5).reshape(N * kb * b, *x0.shape[1:])
            eps_i = eps_i.to(device=net.device, dtype=torch.float64).contiguous()
        else:
            eps_i = torch.randn(N * kb * b, *x0.shape[1:], dtype=torch.float64, device=net.device)
        getu8 = not getattr(params.scorer, 'fused_sums', False)
        x_cand, u8, sums = stepper.step(beams, eps_i, i, want_x_next=True, want_u8=getu8, want_sums=not getu8)
        if getu8:
            s = params.scorer(u8, labels_rows, torch.zeros(u8.shape[0], device=u8.device))
            s = torch.as_tensor(s).to(device=net.device, dtype=torch.float32)
        else:
            s = params.scorer.score_from_sums(sums, C, HW)
        rec.scored_candidates += N * kb * b
        flat = s.reshape(N, kb, b).permute(2, 1, 0).reshape(b, kb * N)     # [b, beam*N + n]
        order = torch.sort(flat, dim=1, descending=True, stable=True).indices[:, :kb]         # [b, k]
        beam_i, n_i = order // N, order % N
        xc = x_cand.reshape(N, kb, b, *x0.shape[1:])
        j = torch.arange(b, device=net.device).unsqueeze(1).expand(b, kb)
        beams = xc[n_i, beam_i, j].permute(1, 0, 2, 3, 4).reshape(kb * b, *x0.shape[1:]).contiguous()
        if record:
            rec.scores.append(flat)
            rec.indices.append(order)
    x_next = beams.reshape(kb, b, *x0.shape[1:])[0].contiguous()          # best beam (:404)
    return x_next, rec


def generate_image_grid(
    network_pkl, dest_path, latents, class_labels,
    seed=0, gridw=8, gridh=8, device=torch.device('cuda'),
    num_steps=18, sigma_min=0.002, sigma_max=80, rho=7,
    S_churn=0, S_min=0, S_max=float('inf'), S_noise=1,
    sampling_method: SamplingMethod = SamplingMethod.NAIVE,
    sampling_params: Optional[Dict[str, Any]] = None,
    precomputed_noise: Optional[Dict[int, torch.Tensor]] = None,
    shard: Optional[Shard] = None, record: bool = False, search_options: Optional[dict] = None,
):
    """Same contract as the reference's generate_image_grid (edm/main.py:47-886).  Returns None
    like the reference unless `record=True`, in which case the SearchRecord is returned.
    `search_options` (not in the reference): keyword options of `eps_greedy_search` for the zero_order / eps_greedy methods,
    e.g. {'escalate': False} or {'kappa': 0.5, 'max_contenders': 4}."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('the B200 path needs a CUDA device; there is no CPU fallback')
    batch_size = gridw * gridh
    torch.manual_seed(seed)
    if sampling_params is None:
        sampling_params = {}
    method_params = SamplingParams(**sampling_params)                     # TypeError on unknown keys, like :64
    print(f'Using sampling method: {sampling_method.name}')
    print(f'Loading network from "{network_pkl}"...')
    net = load_network(network_pkl, device)
    latents = latents.to(device)
    if class_labels is not None:
        class_labels = class_labels.to(device)
    table = StepTable(net, device, num_steps, sigma_min, sigma_max, rho, S_churn, S_min, S_max, S_noise)
    if sampling_method == SamplingMethod.REJECTION_SAMPLING:
        x_next, rec = rejection_search(net, latents, class_labels, method_params, table,
                                       precomputed_noise=precomputed_noise, record=record)
    elif sampling_method == SamplingMethod.BEAM_SEARCH:
        x_next, rec = beam_search(net, latents, class_labels, method_params, table,
                                  precomputed_noise=precomputed_noise, record=record)
    elif sampling_method == SamplingMethod.MCTS:
        x_next, rec = mcts_search(net, latents, class_labels, method_params, table, precomputed_noise=precomputed_noise,
                                  record=record)
    elif sampling_method in (SamplingMethod.ZERO_ORDER, SamplingMethod.EPS_GREEDY):
        print(f"Zero-Order parameters: lambda={method_params.lambda_param}, N={method_params.N}, "
              f"K={method_params.K}, eps={method_params.eps}")
        x_next, rec = eps_greedy_search(net, latents, class_labels, method_params, table,
                                        precomputed_noise=precomputed_noise, shard=shard, record=record,
                                        **(search_options or {}))
    else:
        x_next, rec = naive_search(net, latents, class_labels, table, record=record)

    image = ops.quantize_u8(x_next.contiguous())                          # :869
    timesteps = torch.zeros(image.shape[0], device=device)
    scores = method_params.scorer(image.clone(), class_labels, timesteps)
    avg_score = torch.as_tensor(scores).float().mean().item()
    print(f'Average score: {avg_score}')
    rec.final_image, rec.final_scores = image, scores
    if dest_path is not None and (shard is None or shard.rank == 0):
        import PIL.Image
        print(f'Saving image grid to "{dest_path}"...')
        grid = image.reshape(gridh, gridw, *image.shape[1:]).permute(0, 3, 1, 4, 2)
        grid = grid.reshape(gridh * net.img_resolution, gridw * net.img_resolution, net.img_channels)
        PIL.Image.fromarray(grid.cpu().numpy(), 'RGB').save(dest_path)
    print('Done.')
    return rec if record else None

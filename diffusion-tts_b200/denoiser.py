"""EDM-preconditioned denoiser on the B200 engine + the fused stochastic-Heun candidate step.

`B200Denoiser` follows the net protocol the reference driver relies on
(edm/main.py:80,84,87,882): `net(x, sigma, class_labels) -> fp32 D_x`, `net.round_sigma`,
`net.img_resolution`, `net.img_channels`.  It wraps EDMPrecond.forward
(edm/training/networks.py:654-668) around `UNetEngine`.

`StepTable` holds every candidate-invariant scalar of edm/main.py:82-96 for all steps,
computed ONCE with the very torch expressions the reference evaluates (same device, same
dtypes) and then read back in a single transfer -- the hot loop itself never syncs.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import ops
from .unet import UNetEngine


def load_network(network_pkl: Any, device) -> 'B200Denoiser':
    """Accepts what the reference accepts (a pickle path/URL with an 'ema' entry,
    edm/main.py:69-70) plus: a `.pt` bundle {'state_dict', 'sigma_data'}, a dict of that form,
    any module-like object exposing `.model.state_dict()`, or an existing B200Denoiser."""
    if isinstance(network_pkl, B200Denoiser):
        return network_pkl
    obj = network_pkl
    if isinstance(network_pkl, (str, bytes)):
        path = network_pkl if isinstance(network_pkl, str) else network_pkl.decode()
        if path.endswith('.pt'):
            obj = torch.load(path, map_location='cpu', weights_only=False)
        else:
            try:
                import dnnlib                                    # reference helper (URL cache); optional
                with dnnlib.util.open_url(path) as f:
                    obj = pickle.load(f)
            except ImportError:
                with open(path, 'rb') as f:
                    obj = pickle.load(f)
    if isinstance(obj, dict) and 'ema' in obj:
        obj = obj['ema']
    if isinstance(obj, dict):
        return B200Denoiser(obj['state_dict'], device=device, sigma_data=obj.get('sigma_data', 0.5),
                            sigma_min=obj.get('sigma_min', 0.0), sigma_max=obj.get('sigma_max', float('inf')))
    # duck-typed reference EDMPrecond (re-created from the pickle's embedded source; never isinstance)
    model = getattr(obj, 'model', obj)
    return B200Denoiser(model.state_dict(), device=device, sigma_data=float(getattr(obj, 'sigma_data', 0.5)),
                        sigma_min=float(getattr(obj, 'sigma_min', 0.0)), sigma_max=float(getattr(obj, 'sigma_max', float('inf'))))


class B200Denoiser:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device='cuda', sigma_data=0.5, sigma_min=0.0,
                 sigma_max=float('inf')):
        self.device = torch.device(device)
        self.engine = UNetEngine(state_dict, device=self.device)
        cfg = self.engine.cfg
        self.img_resolution, self.img_channels, self.label_dim = cfg.img_resolution, cfg.in_channels, cfg.label_dim
        self.sigma_data, self.sigma_min, self.sigma_max = sigma_data, sigma_min, sigma_max
        self._state_dict = state_dict            # a reference only: the precise engine packs its weights lazily
        self._precise = None

    @property
    def supports_precise(self) -> bool:
        """The fp32-faithful re-scoring engine (precise.py) implements head_dim-64 attention (ADM) only."""
        cfg = self.engine.cfg
        return all((not b.attention) or b.cout // b.heads == 64 for b in cfg.enc + cfg.dec if b.kind == 'block')

    @property
    def precise_engine(self):
        """Split-fp16 twin of `engine` for near-tie contenders (built on first use)."""
        if self._precise is None:
            from .precise import PreciseUNetEngine
            self._precise = PreciseUNetEngine(self.engine, self._state_dict)
        return self._precise

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError('B200Denoiser is bound to its CUDA device at construction')
        return self

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)

    def precond(self, sigma: torch.Tensor):
        """c_skip, c_out, c_in, c_noise as fp32 tensors (networks.py:656-663)."""
        sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.log() / 4
        return c_skip, c_out, c_in, c_noise

    @torch.no_grad()
    def __call__(self, x, sigma, class_labels=None, **_unused):
        """Generic protocol entry (any batch, one sigma per call or per sample shared by b_emb)."""
        if not x.is_cuda:
            raise RuntimeError('B200Denoiser: input must live on the CUDA device (no CPU fallback)')
        x = x.to(torch.float32)
        sigma = torch.as_tensor(sigma, device=x.device)
        c_skip, c_out, c_in, c_noise = self.precond(sigma)
        labels = None
        b_emb = 1
        if self.label_dim:
            labels = (torch.zeros([1, self.label_dim], device=x.device) if class_labels is None
                      else class_labels.to(torch.float32).reshape(-1, self.label_dim))
            b_emb = self._distinct_rows(labels, x.shape[0])
            labels = labels[:b_emb]
        if c_noise.numel() > 1:
            raise NotImplementedError('per-sample sigma is not used by the search path (edm/main.py:87,92)')
        F_x = self.engine.forward((c_in * x).contiguous(), c_noise.flatten(), labels, b_emb=b_emb)
        return c_skip * x + c_out * F_x.to(torch.float32)

    @staticmethod
    def _distinct_rows(labels: torch.Tensor, B: int) -> int:
        """Labels of a candidate batch are `class_labels.repeat(N,1)` (edm/main.py:806): find the period."""
        n = labels.shape[0]
        if n == 1:
            return 1
        if n != B:
            raise ValueError('class_labels must have 1 or batch rows')
        first = labels[0]
        same = (labels == first).all(dim=1)
        idx = torch.nonzero(same[1:]).flatten()
        period = int(idx[0].item()) + 1 if idx.numel() else n
        if n % period or not torch.equal(labels, labels[:period].repeat(n // period, 1)):
            return n
        return period


@dataclass
class StepCoef:
    t_cur: float
    t_next: float
    t_hat: float
    s: float            # sqrt(t_hat^2 - t_cur^2) * S_noise
    dt: float           # t_next - t_hat
    c_skip1: float
    c_out1: float
    c_in1: float
    c_skip2: float
    c_out2: float
    c_in2: float
    last: bool


class StepTable:
    """Schedule + per-step scalars of edm/main.py:78-96, evaluated with the reference's own
    torch expressions on `device`, transferred once."""

    def __init__(self, net: B200Denoiser, device, num_steps=18, sigma_min=0.002, sigma_max=80, rho=7, S_churn=0,
                 S_min=0, S_max=float('inf'), S_noise=1):
        step_indices = torch.arange(num_steps, dtype=torch.float64, device=device)
        t_steps = (sigma_max ** (1 / rho) + step_indices / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
        t_steps = torch.cat([net.round_sigma(t_steps), torch.zeros_like(t_steps[:1])])
        self.t_steps = t_steps
        t_host = t_steps.cpu()
        rows, cn = [], []
        for i in range(num_steps):
            t_cur, t_next = t_steps[i], t_steps[i + 1]
            gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= float(t_host[i]) <= S_max else 0
            t_hat = net.round_sigma(t_cur + gamma * t_cur)
            s = (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise
            dt = t_next - t_hat
            p1 = net.precond(t_hat)
            # t_next == 0 on the last step: the second evaluation is skipped (edm/main.py:91)
            p2 = net.precond(t_next if i < num_steps - 1 else t_hat)
            rows.append(torch.stack([t_cur, t_next, t_hat, s, dt] + [p.flatten()[0].to(torch.float64) for p in p1[:3]] +
                                    [p.flatten()[0].to(torch.float64) for p in p2[:3]]))
            cn.append(torch.stack([p1[3].flatten()[0], p2[3].flatten()[0]]))
        tab = torch.stack(rows).cpu().tolist()                       # the one device->host transfer
        self.c_noise = torch.stack(cn)                               # [steps, 2] fp32, stays on device
        self.emb = torch.stack([net.engine.positional_embedding(self.c_noise[:, j]) for j in range(2)], dim=1)
        self.steps: List[StepCoef] = []
        for i, r in enumerate(tab):
            self.steps.append(StepCoef(t_cur=r[0], t_next=r[1], t_hat=r[2], s=r[3], dt=r[4], c_skip1=r[5], c_out1=r[6],
                                       c_in1=r[7], c_skip2=r[8], c_out2=r[9], c_in2=r[10], last=i == num_steps - 1))
        self.num_steps = num_steps


class HeunStepper:
    """edm/main.py:82-96 for a batch of R = N*b rows sharing b images' (x_cur, label)."""

    def __init__(self, net: B200Denoiser, table: StepTable, class_labels: Optional[torch.Tensor]):
        self.net, self.table = net, table
        self.labels = None
        if net.label_dim:
            self.labels = (class_labels.to(torch.float32).reshape(-1, net.label_dim) if class_labels is not None
                           else torch.zeros([1, net.label_dim], device=net.device))

    def _forward(self, fp, i: int, which: int):
        fp.emb_in.copy_(self.table.emb[i, which].unsqueeze(0).expand(fp.b_emb, -1))
        fp.plan.run()
        return fp.out

    def _plan(self, R: int, b: int, precise: bool = False, row_images: Optional[torch.Tensor] = None):
        eng = self.net.precise_engine if precise else self.net.engine
        if row_images is not None:
            # arbitrary rows (near-tie contenders of several images): one embedding row per batch row
            fp = eng.plan(R, R)
            if self.labels is not None:
                fp.labels.copy_(self.labels.expand(R, -1) if self.labels.shape[0] == 1 else self.labels[row_images])
            fp._labels_src = None
            return fp
        fp = eng.plan(R, b)
        if self.labels is not None and getattr(fp, '_labels_src', None) is not self.labels:
            fp.labels.copy_(self.labels.expand(fp.b_emb, -1) if self.labels.shape[0] == 1 else self.labels)
            fp._labels_src = self.labels
        return fp

    @torch.no_grad()
    def step(self, x_cur: torch.Tensor, eps: torch.Tensor, i: int, *, want_x_next=True, want_u8=False,
             want_sums=False, precise: bool = False, row_images: Optional[torch.Tensor] = None):
        """x_cur [b,C,H,W] fp64 (shared), eps [R,C,H,W] fp64.  Returns (x_next, x0_u8, chan_sums).

        `precise`: evaluate the network with the split-fp16 (fp32-faithful) engine instead of the bf16 one.
        `row_images` (int64 [R]): row r belongs to image row_images[r] instead of r % b (contender subsets)."""
        c = self.table.steps[i]
        R, b = eps.shape[0], x_cur.shape[0]
        if row_images is not None:
            x_cur = x_cur.index_select(0, row_images).contiguous()
        fp = self._plan(R, b, precise, row_images)
        x_hat, _ = ops.heun_pre(x_cur, eps, c.s, c.c_in1, net_in=fp.x_in)
        F1 = self._forward(fp, i, 0)
        if c.last:
            return ops.heun_post(x_hat, F1, None, c.c_skip1, c.c_out1, c.t_hat, c.dt, want_x_next=want_x_next,
                                 want_u8=want_u8, want_sums=want_sums)
        F1 = F1.clone()
        ops.heun_mid(x_hat, F1, c.c_skip1, c.c_out1, c.t_hat, c.dt, c.c_in2, net_in2=fp.x_in)
        F2 = self._forward(fp, i, 1)
        return ops.heun_post(x_hat, F1, F2, c.c_skip1, c.c_out1, c.t_hat, c.dt, c.c_skip2, c.c_out2, c.t_next,
                             want_x_next=want_x_next, want_u8=want_u8, want_sums=want_sums)

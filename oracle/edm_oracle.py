"""CPU oracle for the candidate-batched noise-search step.  TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference's
algorithm for the hot path.  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import it.  The product (`diffusion-tts_b200/`) never imports anything from `oracle/`.

Pinning: the reference ships no golden vectors or tests for this path (SURVEY.md §4,
§8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in
the build container by `oracle/make_golden.py` (which imports /root/reference) and
committed under `tests/golden/`.  `tests/test_oracle_golden.py` checks every function
here against those fixtures.

Each function cites the reference file:line it restates (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Network structure (edm/training/networks.py:372-433 DhariwalUNet.__init__,
#                    edm/training/networks.py:229-318 SongUNet.__init__)
# --------------------------------------------------------------------------------------


@dataclass
class BlockSpec:
    """One UNetBlock (edm/training/networks.py:134-164) or a bare conv/norm leaf."""
    name: str                 # state-dict prefix, e.g. 'enc.64x64_block0'
    kind: str                 # 'conv' | 'block' | 'aux_norm' | 'aux_conv'
    cin: int
    cout: int
    res: int                  # OUTPUT resolution of the block
    up: bool = False
    down: bool = False
    attention: bool = False
    num_heads: int = 0
    skip_conv: bool = False   # 1x1 skip conv present (networks.py:157-160)
    has_skip: bool = False    # skip module present at all (kernel 0 = pure resample)


@dataclass
class UNetSpec:
    model_type: str
    img_resolution: int
    in_channels: int
    out_channels: int
    label_dim: int
    model_channels: int
    channel_mult: List[int]
    channel_mult_emb: int
    num_blocks: int
    attn_resolutions: List[int]
    emb_channels: int
    noise_channels: int
    skip_scale: float
    eps: float
    adaptive_scale: bool
    enc: List[BlockSpec] = field(default_factory=list)
    dec: List[BlockSpec] = field(default_factory=list)


def build_unet_spec(model_type: str, img_resolution: int, in_channels: int, out_channels: int,
                    label_dim: int = 0, model_channels: Optional[int] = None,
                    channel_mult: Optional[Sequence[int]] = None, channel_mult_emb: int = 4,
                    num_blocks: Optional[int] = None, attn_resolutions: Optional[Sequence[int]] = None,
                    **_unused) -> UNetSpec:
    """Derive the block list the reference constructors build.

    DhariwalUNet: networks.py:395-433.  SongUNet (DDPM++ flavour only: positional
    embedding, standard encoder/decoder, [1,1] filter): networks.py:270-318.
    """
    adm = model_type == 'DhariwalUNet'
    if not adm and model_type != 'SongUNet':
        raise ValueError(f'unsupported model_type {model_type!r}')
    model_channels = model_channels if model_channels is not None else (192 if adm else 128)
    channel_mult = list(channel_mult if channel_mult is not None else ([1, 2, 3, 4] if adm else [1, 2, 2, 2]))
    num_blocks = num_blocks if num_blocks is not None else (3 if adm else 4)
    attn_resolutions = list(attn_resolutions if attn_resolutions is not None else ([32, 16, 8] if adm else [16]))
    spec = UNetSpec(model_type=model_type, img_resolution=img_resolution, in_channels=in_channels,
                    out_channels=out_channels, label_dim=label_dim, model_channels=model_channels,
                    channel_mult=channel_mult, channel_mult_emb=channel_mult_emb, num_blocks=num_blocks,
                    attn_resolutions=attn_resolutions, emb_channels=model_channels * channel_mult_emb,
                    noise_channels=model_channels,
                    skip_scale=1.0 if adm else math.sqrt(0.5), eps=1e-5 if adm else 1e-6,
                    adaptive_scale=adm)

    def heads(cout: int, attention: bool) -> int:
        if not attention:
            return 0
        return cout // 64 if adm else 1          # channels_per_head=64 (:393) vs num_heads=1 (:263)

    def block(prefix, cin, cout, res, up=False, down=False, attention=False) -> BlockSpec:
        has_skip = (cout != cin) or up or down
        # ADM: kernel = 1 iff channels change (resample_proj False);  DDPM++: resample_proj=True (:158)
        skip_conv = has_skip and ((cout != cin) or (not adm))
        return BlockSpec(name=prefix, kind='block', cin=cin, cout=cout, res=res, up=up, down=down,
                         attention=attention, num_heads=heads(cout, attention),
                         skip_conv=skip_conv, has_skip=has_skip)

    cout = in_channels
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            cin, cout = cout, (model_channels * mult if adm else model_channels)
            spec.enc.append(BlockSpec(name=f'enc.{res}x{res}_conv', kind='conv', cin=cin, cout=cout, res=res))
        else:
            spec.enc.append(block(f'enc.{res}x{res}_down', cout, cout, res, down=True))
        for idx in range(num_blocks):
            cin, cout = cout, model_channels * mult
            spec.enc.append(block(f'enc.{res}x{res}_block{idx}', cin, cout, res, attention=res in attn_resolutions))
    skips = [b.cout for b in spec.enc]
    for level, mult in reversed(list(enumerate(channel_mult))):
        res = img_resolution >> level
        if level == len(channel_mult) - 1:
            spec.dec.append(block(f'dec.{res}x{res}_in0', cout, cout, res, attention=True))
            spec.dec.append(block(f'dec.{res}x{res}_in1', cout, cout, res))
        else:
            spec.dec.append(block(f'dec.{res}x{res}_up', cout, cout, res, up=True))
        for idx in range(num_blocks + 1):
            cin, cout = cout + skips.pop(), model_channels * mult
            attn = (res in attn_resolutions) if adm else (idx == num_blocks and res in attn_resolutions)
            spec.dec.append(block(f'dec.{res}x{res}_block{idx}', cin, cout, res, attention=attn))
        if not adm and level == 0:
            spec.dec.append(BlockSpec(name=f'dec.{res}x{res}_aux_norm', kind='aux_norm', cin=cout, cout=cout, res=res))
            spec.dec.append(BlockSpec(name=f'dec.{res}x{res}_aux_conv', kind='aux_conv', cin=cout, cout=out_channels, res=res))
    return spec


def unet_param_shapes(spec: UNetSpec) -> Dict[str, Tuple[int, ...]]:
    """Names/shapes of every learnable tensor of the U-Net (no 'model.' prefix), in a
    fixed order.  Mirrors what the reference modules register (networks.py:30-37,
    49-66, 96-102, 150-164, 398-433)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    adm = spec.model_type == 'DhariwalUNet'
    E, C = spec.emb_channels, spec.noise_channels
    if not adm and spec.label_dim:
        shapes['map_label.weight'] = (C, spec.label_dim)
        shapes['map_label.bias'] = (C,)
    shapes['map_layer0.weight'] = (E, C)
    shapes['map_layer0.bias'] = (E,)
    shapes['map_layer1.weight'] = (E, E)
    shapes['map_layer1.bias'] = (E,)
    if adm and spec.label_dim:
        shapes['map_label.weight'] = (E, spec.label_dim)       # bias=False (:402)

    def conv(prefix, cin, cout, k):
        shapes[f'{prefix}.weight'] = (cout, cin, k, k)
        shapes[f'{prefix}.bias'] = (cout,)

    def norm(prefix, c):
        shapes[f'{prefix}.weight'] = (c,)
        shapes[f'{prefix}.bias'] = (c,)

    for b in spec.enc + spec.dec:
        if b.kind == 'conv' or b.kind == 'aux_conv':
            conv(b.name, b.cin, b.cout, 3)
        elif b.kind == 'aux_norm':
            norm(b.name, b.cin)
        else:
            norm(f'{b.name}.norm0', b.cin)
            conv(f'{b.name}.conv0', b.cin, b.cout, 3)
            shapes[f'{b.name}.affine.weight'] = (b.cout * (2 if spec.adaptive_scale else 1), E)
            shapes[f'{b.name}.affine.bias'] = (b.cout * (2 if spec.adaptive_scale else 1),)
            norm(f'{b.name}.norm1', b.cout)
            conv(f'{b.name}.conv1', b.cout, b.cout, 3)
            if b.skip_conv:
                conv(f'{b.name}.skip', b.cin, b.cout, 1)
            if b.attention:
                norm(f'{b.name}.norm2', b.cout)
                conv(f'{b.name}.qkv', b.cout, b.cout * 3, 1)
                conv(f'{b.name}.proj', b.cout, b.cout, 1)
    if adm:
        norm('out_norm', spec.dec[-1].cout)
        conv('out_conv', spec.dec[-1].cout, spec.out_channels, 3)
    return shapes


def seeded_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic non-degenerate weights for parity runs.

    The reference zero-initialises conv1/proj/out_conv (networks.py:392-393), which would
    make a random-init net compute F_x == 0 (SURVEY.md §7 hard part 2); parity fixtures
    therefore draw EVERY tensor from N(0, 1/fan_in) (biases N(0, 0.1^2), norm weights
    1 + N(0, 0.1^2)) from one seeded CPU generator, in sorted-name order, and feed the
    same tensors to the reference, the oracle and the CUDA engine."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 1:
            t = torch.randn(shp, generator=g) * 0.1
            if name.endswith('weight'):            # GroupNorm gain
                t = t + 1.0
        else:
            fan_in = int(np.prod(shp[1:]))
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        out[name] = t
    return out


# --------------------------------------------------------------------------------------
# U-Net forward (edm/training/networks.py:166-187, 200-206, 320-363, 435-461)
# --------------------------------------------------------------------------------------


def positional_embedding(x: torch.Tensor, num_channels: int, endpoint: bool, max_positions: int = 10000) -> torch.Tensor:
    """networks.py:200-206."""
    half = num_channels // 2
    freqs = torch.arange(0, half, dtype=torch.float32, device=x.device)
    freqs = freqs / (half - (1 if endpoint else 0))
    freqs = (1 / max_positions) ** freqs
    ang = torch.outer(x, freqs.to(x.dtype))
    return torch.cat([ang.cos(), ang.sin()], dim=1)


def _linear(sd, prefix, x):
    y = x @ sd[f'{prefix}.weight'].t()
    if f'{prefix}.bias' in sd:
        y = y + sd[f'{prefix}.bias']
    return y


def _gn(sd, prefix, x, eps):
    c = x.shape[1]
    return F.group_norm(x, num_groups=min(32, c // 4), weight=sd[f'{prefix}.weight'], bias=sd[f'{prefix}.bias'], eps=eps)


def _resample(x, up, down):
    """[1,1] resample filter (networks.py:64-65, 82-85): up = nearest x2, down = 2x2 mean."""
    if up:
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    if down:
        return F.avg_pool2d(x, 2)
    return x


def _conv(sd, prefix, x, k, up=False, down=False):
    x = _resample(x, up, down)
    w = sd.get(f'{prefix}.weight')
    if w is not None:
        x = F.conv2d(x, w, padding=k // 2)
        x = x + sd[f'{prefix}.bias'].reshape(1, -1, 1, 1)
    return x


def embedding(spec: UNetSpec, sd, noise_labels: torch.Tensor, class_labels: Optional[torch.Tensor]) -> torch.Tensor:
    """Mapping network: networks.py:437-447 (ADM), :322-332 (DDPM++)."""
    if spec.model_type == 'DhariwalUNet':
        emb = positional_embedding(noise_labels, spec.noise_channels, endpoint=False)
        emb = F.silu(_linear(sd, 'map_layer0', emb))
        emb = _linear(sd, 'map_layer1', emb)
        if spec.label_dim:
            emb = emb + class_labels @ sd['map_label.weight'].t()
        return F.silu(emb)
    emb = positional_embedding(noise_labels, spec.noise_channels, endpoint=True)
    emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)
    if spec.label_dim:
        emb = emb + _linear(sd, 'map_label', class_labels * math.sqrt(spec.label_dim))
    emb = F.silu(_linear(sd, 'map_layer0', emb))
    return F.silu(_linear(sd, 'map_layer1', emb))


def unet_block(spec: UNetSpec, b: BlockSpec, sd, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """UNetBlock.forward: networks.py:166-187."""
    p = b.name
    orig = x
    x = _conv(sd, f'{p}.conv0', F.silu(_gn(sd, f'{p}.norm0', x, spec.eps)), 3, up=b.up, down=b.down)
    params = _linear(sd, f'{p}.affine', emb)[:, :, None, None]
    if spec.adaptive_scale:
        scale, shift = params.chunk(2, dim=1)
        x = F.silu(torch.addcmul(shift, _gn(sd, f'{p}.norm1', x, spec.eps), scale + 1))
    else:
        x = F.silu(_gn(sd, f'{p}.norm1', x + params, spec.eps))
    x = _conv(sd, f'{p}.conv1', x, 3)
    if b.has_skip:
        x = x + _conv(sd, f'{p}.skip', orig, 1, up=b.up, down=b.down)
    else:
        x = x + orig
    x = x * spec.skip_scale
    if b.attention:
        B, C, H, W = x.shape
        h = b.num_heads
        qkv = _conv(sd, f'{p}.qkv', _gn(sd, f'{p}.norm2', x, spec.eps), 1)
        q, k, v = qkv.reshape(B * h, C // h, 3, H * W).unbind(2)
        w = torch.einsum('ncq,nck->nqk', q, k / math.sqrt(k.shape[1])).softmax(dim=2)   # :116
        a = torch.einsum('nqk,nck->ncq', w, v)
        x = _conv(sd, f'{p}.proj', a.reshape(B, C, H, W), 1) + x
        x = x * spec.skip_scale
    return x


def unet_forward(spec: UNetSpec, sd, x: torch.Tensor, noise_labels: torch.Tensor,
                 class_labels: Optional[torch.Tensor]) -> torch.Tensor:
    """DhariwalUNet.forward networks.py:435-461 / SongUNet.forward networks.py:320-363
    (standard encoder/decoder only)."""
    emb = embedding(spec, sd, noise_labels, class_labels)
    skips = []
    for b in spec.enc:
        x = _conv(sd, b.name, x, 3) if b.kind == 'conv' else unet_block(spec, b, sd, x, emb)
        skips.append(x)
    aux = None
    tmp = None
    for b in spec.dec:
        if b.kind == 'aux_norm':
            tmp = _gn(sd, b.name, x, 1e-6)
        elif b.kind == 'aux_conv':
            aux = _conv(sd, b.name, F.silu(tmp), 3)
        else:
            if x.shape[1] != b.cin:
                x = torch.cat([x, skips.pop()], dim=1)
            x = unet_block(spec, b, sd, x, emb)
    if spec.model_type == 'DhariwalUNet':
        return _conv(sd, 'out_conv', F.silu(_gn(sd, 'out_norm', x, 1e-5)), 3)
    return aux


# --------------------------------------------------------------------------------------
# EDM preconditioning (edm/training/networks.py:654-668)
# --------------------------------------------------------------------------------------


def precond_coeffs(sigma: torch.Tensor, sigma_data: float = 0.5):
    """c_skip, c_out, c_in, c_noise as fp32 tensor ops on the fp32-rounded sigma (:656-663)."""
    sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.log() / 4
    return c_skip, c_out, c_in, c_noise


class OracleNet:
    """Callable with the net protocol the driver uses (edm/main.py:80,84,87,882):
    net(x, sigma, class_labels) -> fp32 D_x; net.round_sigma; img_resolution; img_channels."""

    def __init__(self, spec: UNetSpec, sd: Dict[str, torch.Tensor], sigma_data: float = 0.5):
        self.spec, self.sd, self.sigma_data = spec, sd, sigma_data
        self.img_resolution, self.img_channels, self.label_dim = spec.img_resolution, spec.in_channels, spec.label_dim

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)

    @torch.no_grad()
    def raw(self, x_in: torch.Tensor, c_noise: torch.Tensor, class_labels) -> torch.Tensor:
        return unet_forward(self.spec, self.sd, x_in, c_noise, class_labels)

    @torch.no_grad()
    def __call__(self, x, sigma, class_labels=None):
        x = x.to(torch.float32)
        sigma = torch.as_tensor(sigma)
        if self.label_dim == 0:
            class_labels = None
        elif class_labels is None:
            class_labels = torch.zeros([1, self.label_dim])
        else:
            class_labels = class_labels.to(torch.float32).reshape(-1, self.label_dim)
        c_skip, c_out, c_in, c_noise = precond_coeffs(sigma, self.sigma_data)
        F_x = self.raw(c_in * x, c_noise.flatten(), class_labels)
        return c_skip * x + c_out * F_x.to(torch.float32)


# --------------------------------------------------------------------------------------
# Sampler (edm/main.py:78-99)
# --------------------------------------------------------------------------------------


def karras_schedule(num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0) -> torch.Tensor:
    """edm/main.py:78-80: fp64 rho-schedule with a trailing 0."""
    idx = torch.arange(num_steps, dtype=torch.float64)
    t = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def churn_gamma(t_cur, num_steps, S_churn, S_min, S_max) -> float:
    """edm/main.py:83."""
    return min(S_churn / num_steps, math.sqrt(2) - 1) if S_min <= float(t_cur) <= S_max else 0.0


def heun_step(net, x_cur, t_cur, t_next, i, eps_i, labels, *, num_steps, S_churn=0.0, S_min=0.0,
              S_max=float('inf'), S_noise=1.0):
    """Stochastic Heun step, edm/main.py:82-96.  fp64 state, net I/O fp32.
    Returns (x_next, denoised) where `denoised` is the LAST network output."""
    gamma = churn_gamma(t_cur, num_steps, S_churn, S_min, S_max)
    t_hat = t_cur + gamma * t_cur
    x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * eps_i
    denoised = net(x_hat, t_hat, labels).to(torch.float64)
    d_cur = (x_hat - denoised) / t_hat
    x_next = x_hat + (t_next - t_hat) * d_cur
    if i < num_steps - 1:
        denoised = net(x_next, t_next, labels).to(torch.float64)
        d_prime = (x_next - denoised) / t_next
        x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next, denoised


# --------------------------------------------------------------------------------------
# Scoring (edm/main.py:825-842, edm/scorers.py:25-54, sd/scorers.py:25-76)
# --------------------------------------------------------------------------------------


def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    """edm/main.py:827: (x*127.5+128).clip(0,255).to(uint8) -- truncating cast."""
    return (x * 127.5 + 128).clip(0, 255).to(torch.uint8)


LUMA = (0.2126, 0.7152, 0.0722)


def brightness_score(images: torch.Tensor) -> torch.Tensor:
    """edm/scorers.py:30-54 for uint8 [M,C,H,W]; non-RGB falls back to a plain mean
    (sd/scorers.py:66-67 uses dim=(1,2,3))."""
    if images.dtype == torch.uint8:
        images = images.float() / 255.0
    if images.size(1) == 3:
        w = torch.tensor(LUMA, device=images.device).view(1, 3, 1, 1)
        lum = (images * w).sum(dim=1).mean(dim=(1, 2))
    else:
        lum = images.mean(dim=(1, 2, 3))
    return torch.clamp(lum, 0.0, 1.0)


def argmax_first(scores: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """edm/main.py:842: torch.argmax == first maximal index (exact N-way ties -> 0)."""
    return scores.argmax(dim=dim)


# --------------------------------------------------------------------------------------
# Candidate construction (edm/main.py:749-800)
# --------------------------------------------------------------------------------------


def candidate_scale_seed(i: int, k: int, n: int) -> float:
    """edm/main.py:776: hash(f"{i}_{k}_{n}") % 1000 / 1000 (salted str hash: reproducible
    only under a fixed PYTHONHASHSEED)."""
    return hash(f"{i}_{k}_{n}") % 1000 / 1000.0


def candidate_scale_fp32(scale_seed: float, lambda_scaled: float) -> torch.Tensor:
    """edm/main.py:779: `torch.ones(shape) * scale_seed * lambda_param` -- an fp32 tensor
    rounded after EACH of the two multiplications (lambda_param is a numpy float64 scalar)."""
    return torch.ones([]) * scale_seed * lambda_scaled


def make_candidates(pivot: torch.Tensor, directions: Sequence[Optional[torch.Tensor]],
                    scales: Sequence[torch.Tensor], fresh: Sequence[Optional[torch.Tensor]]) -> torch.Tensor:
    """N candidates around `pivot` [b,C,H,W] (edm/main.py:749-800).
    directions[n] (un-normalised) is used when fresh[n] is None; scales[n] is the 0-d fp32
    tensor of candidate_scale_fp32 (broadcast like the reference's [b,1,1,1] tensor, :779)
    multiplying the fp64 unit direction."""
    out = []
    for n in range(len(scales)):
        if fresh[n] is not None:
            out.append(fresh[n])
            continue
        d = directions[n]
        d = d / torch.norm(d, p=2, dim=tuple(range(1, d.dim())), keepdim=True)
        out.append(pivot + scales[n] * d)
    return torch.cat(out, dim=0)


# --------------------------------------------------------------------------------------
# Search drivers (edm/main.py:101-137 rejection, :714-860 eps_greedy/zero_order, :862-866 naive)
# --------------------------------------------------------------------------------------


@dataclass
class SearchRecord:
    scores: List[torch.Tensor] = field(default_factory=list)       # per (i,k): [N,b]
    indices: List[torch.Tensor] = field(default_factory=list)      # per (i,k): [b]
    pivots: List[torch.Tensor] = field(default_factory=list)       # per i: committed noise [b,C,H,W]
    x_steps: List[torch.Tensor] = field(default_factory=list)      # per i: committed x_next fp64
    final_image: Optional[torch.Tensor] = None
    final_scores: Optional[torch.Tensor] = None


def eps_greedy_search(net, latents, class_labels, scorer: Callable, *, N, K, lambda_param, eps,
                      noise: Dict, num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0,
                      S_churn=0.0, S_min=0.0, S_max=float('inf'), S_noise=1.0,
                      bernoulli: Optional[Callable[[int, int, int], bool]] = None,
                      scale_fn: Optional[Callable[[int, int, int], float]] = None,
                      teacher: Optional[SearchRecord] = None) -> SearchRecord:
    """ZERO_ORDER == EPS_GREEDY branch, edm/main.py:714-860.

    `noise` follows the reference's precomputed_noise protocol: 'pivot_{i}' [b,C,H,W],
    i -> [b,K,N,C,H,W] directions, 'fresh_{i}_{k}_{n}' [b,C,H,W].  `bernoulli(i,k,n)`
    returns True for the perturbation branch (reference: torch.rand(1) < 1-eps, :751);
    default: eps == 0 -> always perturb, eps == 1 -> always fresh.  `scale_fn(i,k,n)`
    returns hash(f"{i}_{k}_{n}") % 1000 / 1000 (default: this process's salted hash()).
    """
    if bernoulli is None:
        assert eps in (0, 0.0, 1, 1.0), 'supply bernoulli() to replay the reference RNG for 0<eps<1'
        bernoulli = lambda i, k, n: eps == 0
    lam = lambda_param * np.sqrt(3 * 64 * 64)                      # :716 (hard-coded 64x64x3)
    t_steps = karras_schedule(num_steps, sigma_min, sigma_max, rho)
    b = latents.shape[0]
    rec = SearchRecord()
    x_next = latents.to(torch.float64) * t_steps[0]
    kw = dict(num_steps=num_steps, S_churn=S_churn, S_min=S_min, S_max=S_max, S_noise=S_noise)
    for i in range(num_steps):
        t_cur, t_next = t_steps[i], t_steps[i + 1]
        x_cur = x_next
        pivot = noise[f'pivot_{i}']
        for k in range(K):
            dirs, fresh, scales = [], [], []
            for n in range(N):
                if bernoulli(i, k, n):
                    dirs.append(noise[i][:, k, n].reshape(pivot.shape))
                    fresh.append(None)
                    seed_ikn = scale_fn(i, k, n) if scale_fn else candidate_scale_seed(i, k, n)
                    scales.append(candidate_scale_fp32(seed_ikn, lam))
                else:
                    dirs.append(None)
                    fresh.append(noise[f'fresh_{i}_{k}_{n}'])
                    scales.append(None)
            cands = make_candidates(pivot, dirs, scales, fresh)                   # [N*b,...]
            labels_exp = class_labels.repeat(N, 1) if class_labels is not None else None
            _, x0 = heun_step(net, x_cur.repeat(N, 1, 1, 1), t_cur, t_next, i, cands, labels_exp, **kw)
            u8 = quantize_u8(x0)
            scores = scorer(u8, labels_exp, torch.zeros(u8.shape[0])).reshape(N, b)
            best = argmax_first(scores, dim=0)
            cb = cands.reshape(N, b, *cands.shape[1:])
            pivot = torch.stack([cb[best[j], j] for j in range(b)])
            rec.scores.append(scores.clone())
            rec.indices.append(best.clone())
        rec.pivots.append(pivot.clone())
        x_next, _ = heun_step(net, x_cur, t_cur, t_next, i, pivot, class_labels, **kw)   # :860
        rec.x_steps.append(x_next.clone())
        if teacher is not None:            # teacher forcing: continue from the teacher's committed state
            x_next = teacher.x_steps[i].clone()
    rec.final_image = quantize_u8(x_next)
    rec.final_scores = scorer(rec.final_image, class_labels, torch.zeros(b))
    return rec


def naive_search(net, latents, class_labels, scorer, *, noise: Sequence[torch.Tensor], num_steps=18,
                 sigma_min=0.002, sigma_max=80.0, rho=7.0, S_churn=0.0, S_min=0.0, S_max=float('inf'),
                 S_noise=1.0) -> SearchRecord:
    """NAIVE branch, edm/main.py:862-866, with the per-step randn supplied in `noise[i]`."""
    t_steps = karras_schedule(num_steps, sigma_min, sigma_max, rho)
    rec = SearchRecord()
    x_next = latents.to(torch.float64) * t_steps[0]
    for i in range(num_steps):
        x_next, _ = heun_step(net, x_next, t_steps[i], t_steps[i + 1], i, noise[i], class_labels,
                              num_steps=num_steps, S_churn=S_churn, S_min=S_min, S_max=S_max, S_noise=S_noise)
        rec.x_steps.append(x_next.clone())
    rec.final_image = quantize_u8(x_next)
    rec.final_scores = scorer(rec.final_image, class_labels, torch.zeros(latents.shape[0]))
    return rec


def rejection_search(net, latents, class_labels, scorer, *, N, noise: Dict[int, torch.Tensor], num_steps=18,
                     sigma_min=0.002, sigma_max=80.0, rho=7.0, S_churn=0.0, S_min=0.0, S_max=float('inf'),
                     S_noise=1.0) -> SearchRecord:
    """REJECTION_SAMPLING branch, edm/main.py:101-137: N whole trajectories per image
    (repeat_interleave layout: row = j*N + n), final-image score, argmax per image.
    noise[i] is [b, maxN, C, H, W]."""
    t_steps = karras_schedule(num_steps, sigma_min, sigma_max, rho)
    b = latents.shape[0]
    rec = SearchRecord()
    x = (latents.to(torch.float64) * t_steps[0]).repeat_interleave(N, dim=0)
    labels = class_labels.repeat_interleave(N, dim=0)
    for i in range(num_steps):
        eps_i = noise[i][:, :N].reshape(b * N, *x.shape[1:])
        x, _ = heun_step(net, x, t_steps[i], t_steps[i + 1], i, eps_i, labels, num_steps=num_steps,
                         S_churn=S_churn, S_min=S_min, S_max=S_max, S_noise=S_noise)
    u8 = quantize_u8(x)
    scores = scorer(u8, labels, torch.zeros(u8.shape[0])).view(b, N)
    best = scores.argmax(dim=1)
    xr = x.view(b, N, *x.shape[1:])
    x_next = torch.stack([xr[j, best[j]] for j in range(b)])
    rec.scores.append(scores.clone())
    rec.indices.append(best.clone())
    rec.x_steps.append(x_next.clone())
    rec.final_image = quantize_u8(x_next)
    rec.final_scores = scorer(rec.final_image, class_labels, torch.zeros(b))
    return rec


# --------------------------------------------------------------------------------------
# MCTS (edm/main.py:405-713) -- SURVEY.md 8 f4
# --------------------------------------------------------------------------------------


class MCTSNode:
    """One tree node: the state x at depth `depth` (= number of steps taken), its children, visit / reward sums."""
    __slots__ = ('x', 'depth', 'children', 'reward', 'visit')

    def __init__(self, x, depth, visit=0):
        self.x, self.depth, self.children, self.reward, self.visit = x, depth, [], 0.0, visit


def mcts_select(root: MCTSNode) -> List[MCTSNode]:
    """Selection (edm/main.py:548-572): descend by UCB1 while the node has children; an unvisited child scores +inf and
    np.argmax takes the FIRST maximum.  Returns the path root..leaf."""
    import numpy as np
    path = [root]
    node = root
    while node.children:
        ucb = []
        for ch in node.children:
            if ch.visit == 0:
                ucb.append(float('inf'))
            else:
                ucb.append(ch.reward / ch.visit + np.sqrt(2 * np.log(node.visit) / ch.visit))
        node = node.children[int(np.argmax(ucb))]
        path.append(node)
    return path


def mcts_search(net, latents, class_labels, scorer: Callable, *, N: int, S: int, noise: Optional[Dict[int, torch.Tensor]] = None,
                num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0, record: Optional[dict] = None, **sampler_kw):
    """SamplingMethod.MCTS (edm/main.py:405-713) for the images of `latents`, in the reference's mini-batches of
    min(2, batch) samples.  b = N children per expansion, S simulations per timestep, run in groups of 16 whose statistics
    are only backed up after the whole group (:521, :664-679).  Per timestep: expand the root if needed (:467-512); every
    simulation selects a leaf (mcts_select), expands it unless it is at the last step (:575-591; all nodes of one depth
    share the b noises of that depth), descends into a np.random child (:594), rolls out DETERMINISTICALLY (zero noise) to
    t = 0 (:617-640) and is scored on the final image (:657-661); the root moves to its visited child with the best mean
    reward, first maximum (:682-700), keeping its subtree.
    RNG: like the reference, the per-depth noises come from torch.randn (fp32) unless `noise[i]` ([1, b, C, H, W]) is
    given, every expansion evaluates one throw-away torch.randn (the eager default of the dict .get at :578), and the
    rollout child comes from np.random.randint -- seed both generators to reproduce a run."""
    import numpy as np
    t_steps = karras_schedule(num_steps, sigma_min, sigma_max, rho)
    x_all = latents.to(torch.float64) * t_steps[0]
    batch = x_all.shape[0]
    b = N
    results = []
    mbs = min(2, batch)
    for mb0 in range(0, batch, mbs):
        mb = min(mb0 + mbs, batch) - mb0
        xb = x_all[mb0:mb0 + mb]
        lb = None if class_labels is None else class_labels[mb0:mb0 + mb]
        lab = lambda s: None if lb is None else lb[s:s + 1]
        depth_noise = {}
        for i in range(num_steps):                                             # :439-447
            if noise is not None and i in noise:
                depth_noise[i] = noise[i].repeat(mb, 1, 1, 1, 1)
            else:
                depth_noise[i] = torch.randn(mb, b, *xb.shape[1:])
        roots = [MCTSNode(xb[s:s + 1].clone(), 0, visit=1) for s in range(mb)]

        def expand(node: MCTSNode, s: int, throwaway: bool):
            i = node.depth
            for n in range(b):
                if throwaway:
                    torch.randn(1, *xb.shape[1:])                              # eager default argument of .get (:578)
                x_child, _ = heun_step(net, node.x, t_steps[i], t_steps[i + 1], i, depth_noise[i][s, n:n + 1], lab(s),
                                       num_steps=num_steps, **sampler_kw)
                node.children.append(MCTSNode(x_child, i + 1))

        for i in range(num_steps):
            need = [s for s in range(mb) if not roots[s].children]             # root expansion, ONE batched step (:467-512)
            if need:
                xs = torch.cat([roots[s].x for s in need for _ in range(b)])
                es = torch.cat([depth_noise[i][s:s + 1, n] for s in need for n in range(b)])
                ls = None if lb is None else torch.cat([lb[s:s + 1] for s in need for _ in range(b)])
                xc, _ = heun_step(net, xs, t_steps[i], t_steps[i + 1], i, es, ls, num_steps=num_steps, **sampler_kw)
                for r, s in enumerate(s for s in need for _ in range(b)):
                    roots[s].children.append(MCTSNode(xc[r:r + 1], i + 1))
            group = min(16, S * mb)
            for g0 in range(0, S * mb, group):
                paths, starts = [], []
                for sim in range(g0, min(g0 + group, S * mb)):
                    s = sim % mb
                    path = mcts_select(roots[s])
                    leaf = path[-1]
                    if leaf.depth < num_steps - 1:                             # :575 (i_traverse < len(t_steps) - 2)
                        expand(leaf, s, throwaway=True)
                        leaf = leaf.children[np.random.randint(0, len(leaf.children))]
                        path.append(leaf)
                    paths.append(path)
                    starts.append((leaf.x.clone(), leaf.depth, s))
                finals = []
                for x, d, s in starts:                                         # deterministic rollout (:617-640)
                    for j in range(d, num_steps):
                        x, _ = heun_step(net, x, t_steps[j], t_steps[j + 1], j, torch.zeros_like(x), lab(s),
                                         num_steps=num_steps, **sampler_kw)
                    finals.append(x)
                img = quantize_u8(torch.cat(finals))
                labs = None if lb is None else torch.cat([lb[s:s + 1] for _, _, s in starts])
                rewards = scorer(img, labs, torch.zeros(img.shape[0]))
                if record is not None:
                    record.setdefault('rewards', []).append(rewards.clone())
                    record.setdefault('depths', []).append([d for _, d, _ in starts])
                for path, r in zip(paths, rewards):                            # backup after the whole group (:664-679)
                    for node in path:
                        node.reward += r.item()
                        node.visit += 1
            for s in range(mb):                                                # :682-700
                best, best_r = None, -float('inf')
                for k, ch in enumerate(roots[s].children):
                    if ch.visit > 0 and ch.reward / ch.visit > best_r:
                        best, best_r = ch, ch.reward / ch.visit
                assert best is not None
                if record is not None:
                    record.setdefault('chosen', []).append(roots[s].children.index(best))
                    record.setdefault('x', []).append(best.x.clone())
                roots[s] = best
        results.extend(r.x for r in roots)
    return torch.cat(results)

"""Scorers with the reference's call protocol (edm/scorers.py:14-23):
`scorer(images: uint8 [M,C,H,W] | list[PIL], class_labels | prompts, timesteps) -> float[M]`,
higher is better.  The B200 implementations run on integer channel sums produced by
hand-written CUDA kernels (a warp-shuffle reduction per candidate), and expose
`score_from_sums` so the search loop can fuse scoring into the sampler epilogue without ever
materialising the uint8 images.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class Scorer(torch.nn.Module):
    """Base class (edm/scorers.py:14-23)."""

    def __init__(self, dtype=torch.float32):
        super().__init__()
        self.dtype = dtype
        self.eval()

    @torch.no_grad()
    def __call__(self, images, prompts, timesteps=None):
        raise NotImplementedError('Subclasses must implement __call__')


class BrightnessScorer(Scorer):
    """Perceived luminance 0.2126 R + 0.7152 G + 0.0722 B averaged over pixels, clamped to
    [0,1] (edm/scorers.py:25-54); non-RGB inputs fall back to the plain mean
    (sd/scorers.py:66-67).  Computed from exact integer sums: differs from the reference's
    fp32 pipeline by at most 1 ulp; identical images always score identically (exact ties)."""

    fused_sums = True      # tells the search loop it can pass heun_post's channel sums directly

    def __init__(self, dtype=torch.float32, device='cuda'):
        super().__init__(dtype)
        self.device = torch.device(device)

    def score_from_sums(self, chan_sums: torch.Tensor, channels: int, hw: int) -> torch.Tensor:
        return ops.brightness_from_sums(chan_sums, channels, hw)

    @torch.no_grad()
    def __call__(self, images, prompts=None, timesteps=None):
        if isinstance(images, list):       # list of PIL images (edm/scorers.py:31-34)
            images = torch.stack([torch.from_numpy(np.array(img)).permute(2, 0, 1) for img in images])
        if not isinstance(images, torch.Tensor) or images.dtype != torch.uint8:
            raise TypeError('B200 BrightnessScorer scores uint8 images (the format the search loop produces, '
                            'edm/main.py:827); got ' + str(getattr(images, 'dtype', type(images))))
        if images.dim() != 4:
            raise ValueError('expected [M,C,H,W] images')
        if not images.is_cuda:
            images = images.to(self.device)
        images = images.contiguous()
        M, C, H, W = images.shape
        if C > 4:
            raise ValueError('at most 4 channels are supported')
        return self.score_from_sums(ops.channel_sums_u8(images), C, H * W).to(self.dtype)

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
    latent_fused = True    # SD beam: the 4-channel plain-mean branch is fused into ddim_x0_score_kernel

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


class CompressibilityScorer(Scorer):
    """`1 - clip((jpeg_bytes - min)/(max - min), 0, 1)` with the byte count of the baseline JPEG that
    `PIL.Image.save(format='JPEG', quality=q)` would write (edm/scorers.py:176-244) -- computed on the GPU
    by a bit-exact port of libjpeg's integer pipeline (csrc/jpeg.cuh), including 0xFF byte stuffing.
    The quantisation/Huffman tables and the header length are read once per (H, W, quality) from a header
    that the installed libjpeg itself writes for an all-zero image (host-side setup, not on the hot path).
    Images must be uint8 RGB [M,3,H,W] with H, W multiples of 16 and <= 64 (the EDM search path)."""

    def __init__(self, quality=80, min_size=0, max_size=3000, dtype=torch.float32, device='cuda'):
        super().__init__(dtype)
        self.quality, self.min_size, self.max_size = quality, min_size, max_size
        self.device = torch.device(device)
        self._tables = {}

    @staticmethod
    def _parse_header(data: bytes):
        zigzag = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                  21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                  60, 61, 54, 47, 55, 62, 63]
        q = np.zeros((2, 64), dtype=np.int32)
        dc_len, dc_code = np.zeros((2, 16), dtype=np.int32), np.zeros((2, 16), dtype=np.int32)
        ac_len, ac_code = np.zeros((2, 256), dtype=np.int32), np.zeros((2, 256), dtype=np.int32)
        i = 2
        while i < len(data):
            marker, seglen = data[i + 1], (data[i + 2] << 8) | data[i + 3]
            body = data[i + 4:i + 2 + seglen]
            if marker == 0xDB:                                   # DQT (zig-zag order in the file)
                j = 0
                while j < len(body):
                    for k in range(64):
                        q[body[j] & 15, zigzag[k]] = body[j + 1 + k]
                    j += 65
            elif marker == 0xC4:                                 # DHT: canonical code assignment
                j = 0
                while j < len(body):
                    tc, th = body[j] >> 4, body[j] & 15
                    bits = body[j + 1:j + 17]
                    vals = body[j + 17:j + 17 + sum(bits)]
                    lens, codes = (dc_len, dc_code) if tc == 0 else (ac_len, ac_code)
                    code, kk = 0, 0
                    for length in range(1, 17):
                        for _ in range(bits[length - 1]):
                            lens[th, vals[kk]], codes[th, vals[kk]] = length, code
                            code += 1
                            kk += 1
                        code <<= 1
                    j += 17 + sum(bits)
            elif marker == 0xDA:                                 # SOS: entropy-coded data starts after it
                header = i + 2 + seglen
                return np.concatenate([q.ravel(), dc_len.ravel(), dc_code.ravel(), ac_len.ravel(), ac_code.ravel(),
                                       np.array([header], dtype=np.int32)])
            i += 2 + seglen
        raise ValueError('no SOS marker in the JPEG header')

    def _get_tables(self, H, W):
        key = (H, W)
        if key not in self._tables:
            import io

            from PIL import Image

            from . import _lib
            buf = io.BytesIO()
            Image.fromarray(np.zeros((H, W, 3), dtype=np.uint8)).save(buf, format='JPEG', quality=self.quality)
            flat = self._parse_header(buf.getvalue())
            assert flat.size * 4 == _lib.lib().b200ns_jpeg_tables_bytes()
            self._tables[key] = torch.from_numpy(flat).to(self.device)
        return self._tables[key]

    @torch.no_grad()
    def sizes_and_scores(self, images: torch.Tensor):
        from . import _lib
        if not isinstance(images, torch.Tensor) or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[1] != 3:
            raise TypeError('B200 CompressibilityScorer scores uint8 RGB [M,3,H,W] images (edm/main.py:827)')
        images = images.to(self.device).contiguous()
        M, _, H, W = images.shape
        tab = self._get_tables(H, W)
        sizes = torch.empty(M, dtype=torch.int32, device=self.device)
        scores = torch.empty(M, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().b200ns_jpeg_size(images.data_ptr(), tab.data_ptr(), M, H, W, float(self.min_size),
                                               float(self.max_size), sizes.data_ptr(), scores.data_ptr(),
                                               _lib.cur_stream()), 'jpeg_size')
        ops.LAUNCHES[0] += 1
        return sizes, scores

    @torch.no_grad()
    def __call__(self, images, prompts=None, timesteps=None):
        return self.sizes_and_scores(images)[1].to(self.dtype)

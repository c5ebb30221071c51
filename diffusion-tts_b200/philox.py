"""Host-side mirror of N consecutive `torch.rand(1, device='cuda')` draws (edm/main.py:751 calls it once per candidate).

The reference decides "perturb the pivot or draw fresh noise" with one `torch.rand(1) < 1 - eps` per candidate: on the
GPU that is 2 tiny kernels per candidate in the stream (1024 launches per round at N=512, ~2.8 ms of serialised launch
latency).  The numbers themselves are a pure function of the CUDA generator's (seed, offset): ATen's
`distribution_elementwise_grid_stride_kernel` gives element 0 of a 1-element tensor the first lane of
Philox4x32-10(key=seed, counter=(offset/4, 0, subsequence=0)) mapped by curand_uniform, and every call advances the
offset by 4.  `rand1_sequence` evaluates that on the host with numpy and advances the generator's offset, so the RNG
stream afterwards is exactly where the reference's N calls would have left it -- no kernel, no synchronisation.

`mirror_ok(device)` checks the emulation once per process against real `torch.rand(1)` calls (restoring the generator
state); the search loop uses the per-call form when it does not hold (another torch version), never silently wrong."""
from __future__ import annotations

import numpy as np
import torch

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr uint32 [n,4], key uint32 [n,2] -> uint32 [n,4] (Random123 / curand Philox4x32-10)."""
    c = ctr.astype(np.uint64)
    k0, k1 = key[:, 0].astype(np.uint64), key[:, 1].astype(np.uint64)
    c0, c1, c2, c3 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    for r in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        if r < 9:
            k0, k1 = (k0 + np.uint64(_W0)) & _MASK, (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _uniform_from_bits(x: np.ndarray) -> np.ndarray:
    """curand_uniform: `x * 2^-32 + 2^-33` in fp32 -- the uint32 is first converted to fp32 (round to nearest even,
    24 bits), the scaling by a power of two is exact, the add rounds once (fused or not) -- then ATen maps 1.0 -> 0.0."""
    xf = x.astype(np.float32).astype(np.float64)
    u = (xf * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)            # the fp64 sum is exact => one rounding, like the GPU
    return np.where(u == np.float32(1.0), np.float32(0.0), u)


def rand1_values(seed: int, offset: int, n: int) -> np.ndarray:
    """Values of n consecutive torch.rand(1) calls when the generator is at (seed, offset); fp32 [n]."""
    offs = (offset // 4) + np.arange(n, dtype=np.uint64)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (offs & _MASK).astype(np.uint32)
    ctr[:, 1] = (offs >> np.uint64(32)).astype(np.uint32)
    key = np.empty((n, 2), dtype=np.uint32)
    key[:, 0], key[:, 1] = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    return _uniform_from_bits(philox4x32_10(ctr, key)[:, 0])


def _generator(device) -> torch.Generator:
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[idx]


def rand1_sequence(device, n: int) -> np.ndarray:
    """The next n torch.rand(1, device=device) values, consuming them from the device's default generator."""
    g = _generator(device)
    off = g.get_offset()
    vals = rand1_values(g.initial_seed(), off, n)
    g.set_offset(off + 4 * n)
    return vals


_OK = {}


def mirror_ok(device) -> bool:
    """One-time self check against the real thing (32 draws; generator state restored afterwards)."""
    key = str(torch.device(device))
    if key not in _OK:
        try:
            g = _generator(device)
            state = g.get_state()
            want = rand1_values(g.initial_seed(), g.get_offset(), 32)
            got = torch.cat([torch.rand(1, device=device) for _ in range(32)]).cpu().numpy()
            end = g.get_offset()
            g.set_state(state)
            _OK[key] = bool(np.array_equal(want, got)) and end == g.get_offset() + 128 and g.get_offset() % 4 == 0
        except Exception:
            _OK[key] = False
    return _OK[key]

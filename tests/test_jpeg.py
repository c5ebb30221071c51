"""Compressibility scorer.  CPU: the numpy restatement of libjpeg's pipeline
(oracle/jpeg_oracle.py) reproduces PIL's real byte counts.  GPU: the CUDA kernel reproduces them too
(bit-exact sizes, hence identical scores and argmax)."""
import numpy as np
import pytest
import torch

from oracle import jpeg_oracle as J


def _images(n=24, size=64, seed=0):
    rng = np.random.default_rng(seed)
    out = [np.zeros((size, size, 3), np.uint8), np.full((size, size, 3), 255, np.uint8),
           rng.integers(0, 256, (size, size, 3), dtype=np.uint8),
           (rng.integers(0, 2, (size, size, 1)) * 255).repeat(3, -1).astype(np.uint8)]
    x = np.linspace(0, 255, size)
    out.append(np.stack([np.tile(x, (size, 1)), np.tile(x[:, None], (1, size)), 128 * np.ones((size, size))], -1).astype(np.uint8))
    while len(out) < n:                                   # smooth random fields at several contrasts (image-like)
        f = rng.normal(0, 1, (size, size, 3))
        for _ in range(int(rng.integers(0, 6))):
            f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 0) + np.roll(f, -1, 1)) / 5
        out.append(np.clip(f * rng.uniform(20, 600) + rng.uniform(40, 200), 0, 255).astype(np.uint8))
    return out


def test_jpeg_oracle_matches_pil_byte_counts():
    for im in _images(10):
        assert J.encode_image(im)[0] == J.pil_size(im)
    for im in _images(4, size=32, seed=1):
        assert J.encode_image(im)[0] == J.pil_size(im)
    assert J.pil_size(np.zeros((64, 64, 3), np.uint8)) == 691          # SURVEY.md 7 hard part 7
    assert J.compressibility_score(None, size=691) == pytest.approx(1 - 691 / 3000)


@pytest.mark.gpu
@pytest.mark.parametrize('size', [64, 32])
def test_jpeg_kernel_matches_pil_exactly(size):
    from diffusion_tts_b200 import build
    build.build()
    from diffusion_tts_b200.scorers import CompressibilityScorer
    ims = _images(24, size=size, seed=size)
    batch = torch.from_numpy(np.stack(ims)).permute(0, 3, 1, 2).contiguous().cuda()
    scorer = CompressibilityScorer(device='cuda')
    sizes, scores = scorer.sizes_and_scores(batch)
    want = [J.pil_size(im) for im in ims]
    assert sizes.cpu().tolist() == want
    ref = torch.tensor([1.0 - min(1.0, max(0.0, (s - 0) / 3000)) for s in want])      # edm/scorers.py:243
    assert torch.equal(scores.cpu(), ref)
    assert torch.equal(scorer(batch, None, None).cpu(), ref)

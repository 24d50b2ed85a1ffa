"""Synthetic inputs for bench.py (both arms): seeded multi-octave images in [0, 1].

The generator is the recipe of SURVEY.md Appendix C (octaves of uniform noise, each bilinearly enlarged to the image
size, amplitudes growing 1.6x towards the coarse octaves, min-max normalised per image).  It lives here rather than in
oracle/ so that the GPU arm's inputs do not come from test infrastructure; tests/test_oracle.py checks that it produces
the same tensor as the oracle's copy.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def octave_images(b: int, s: int, g: torch.Generator) -> torch.Tensor:
    x = torch.zeros(b, 3, s, s)
    amp, tot, r = 1.0, 0.0, s
    while r >= 5:
        x += amp * F.interpolate(torch.rand(b, 3, r, r, generator=g), size=(s, s), mode="bilinear", align_corners=False)
        tot += amp
        r //= 2
        amp *= 1.6
    x /= tot
    lo, hi = x.amin(dim=(1, 2, 3), keepdim=True), x.amax(dim=(1, 2, 3), keepdim=True)
    return (x - lo) / (hi - lo)


def make_inputs(batch: int, img: int = 640, seed: int = 7) -> torch.Tensor:
    """fp32 [batch, 3, img, img] in [0, 1] (matches the /255 preprocessing).  With the calibrated weights ~1-5 % of the
    anchors pass conf 0.25, so NMS does real work.  Eight distinct images are generated and repeated with a per-image
    brightness ramp (keeps values in [0, 1], de-duplicates the repeats)."""
    g = torch.Generator().manual_seed(seed)
    base = octave_images(min(batch, 8), img, g)
    reps = -(-batch // base.shape[0])
    x = base.repeat(reps, 1, 1, 1)[:batch].clone()
    x *= torch.linspace(0.85, 1.0, batch).view(-1, 1, 1, 1)
    return x

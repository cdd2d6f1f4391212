"""Frame sources for tests and the benchmark: 8-bit luma ``.yuv`` reading
(layout of ``src/common/utils.c:61-73``: the first W*H bytes of the file) and the
deterministic synthetic stand-ins of SURVEY.md section 8(d) for the Beauty /
Jockey frames that are absent from the reference checkout
(``.MISSING_LARGE_BLOBS``).  Every generator is seeded and pure numpy, so the
same frames come out in the build container and on the GPU box.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

# the reference's frames/ directory (ForemanYF{1,2,4}.yuv, 352x288 luma), shipped with the package
FRAMES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def read_yuv_luma(path: str, width: int, height: int) -> np.ndarray:
    """First luma plane of an 8-bit .yuv file as (H, W) uint8."""
    with open(path, "rb") as f:
        buf = f.read(width * height)
    if len(buf) != width * height:
        raise IOError(f"{path}: short read")
    return np.frombuffer(buf, np.uint8).reshape(height, width).copy()


def foreman(idx: int) -> np.ndarray:
    """Shipped Foreman CIF luma frame YF<idx> (352x288), idx in {1, 2, 4}."""
    return read_yuv_luma(os.path.join(FRAMES_DIR, f"ForemanYF{idx}.yuv"), 352, 288)


def tiled_frames(width: int, height: int, cur_idx: int = 2, ref_idx: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """'tiled-Foreman': real texture and real motion at any size."""
    def tile(a):
        ry, rx = -(-height // a.shape[0]), -(-width // a.shape[1])
        return np.ascontiguousarray(np.tile(a, (ry, rx))[:height, :width])
    return tile(foreman(cur_idx)), tile(foreman(ref_idx))


def shifted_noise_pair(width: int, height: int, seed: int = 1234, shift=(5, -3), cell: int = 8,
                       sigma: float = 4.0) -> Tuple[np.ndarray, np.ndarray]:
    """'shifted-noise' (SURVEY 8d-ii): blocky uniform texture; the current frame is
    the reference shifted by `shift` (dx, dy) plus Gaussian noise, clipped."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gh, gw = -(-height // cell) + 2, -(-width // cell) + 2
    coarse = rng.integers(0, 256, size=(gh, gw), dtype=np.int32)
    ref = np.kron(coarse, np.ones((cell, cell), np.int32))[:height + cell, :width + cell]
    dx, dy = shift
    cur = np.roll(ref, (dy, dx), axis=(0, 1)).astype(np.float64)
    cur = cur + rng.normal(0.0, sigma, size=cur.shape)
    cur = np.clip(np.rint(cur), 0, 255).astype(np.uint8)[:height, :width]
    return np.ascontiguousarray(cur), np.ascontiguousarray(ref[:height, :width].astype(np.uint8))


def constant_pair(width: int, height: int, value: int = 128):
    """All candidates tie: exercises the first-minimum tie-break (main.c:56)."""
    a = np.full((height, width), value, np.uint8)
    return a, a.copy()


def random_pair(width: int, height: int, seed: int = 7):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.integers(0, 256, (height, width), dtype=np.uint8),
            rng.integers(0, 256, (height, width), dtype=np.uint8))


def checker_pair(width: int, height: int, cell: int = 1):
    """Saturated 0/255 checkerboards in opposite phase: maximal SSD."""
    yy, xx = np.mgrid[0:height, 0:width]
    a = (((yy // cell) + (xx // cell)) & 1).astype(np.uint8) * 255
    return a, (255 - a).astype(np.uint8)


def far_pair(width: int, height: int, seed: int = 11):
    """Dark current frame vs bright reference: every SSD of a block with more than
    258 pixels exceeds 2^24, where the reference's float accumulation (main.c:19-26)
    starts to round -- the adversarial case for score exactness."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.integers(0, 41, (height, width), dtype=np.uint8),
            rng.integers(215, 256, (height, width), dtype=np.uint8))


def inverted_pair(width: int, height: int, seed: int = 21, period: float = 37.0):
    """Smooth texture vs its negative: every candidate of most blocks has a negative
    cross-covariance, so no SSIM score exceeds 0 -- the case in which the reference's
    SSIM scan never sets a motion vector (ssim.c:88-103)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:height, 0:width]
    ph = rng.uniform(0, 6.28, 3)
    t = (np.sin(xx / period + ph[0]) + np.sin(yy / (0.8 * period) + ph[1]) + np.sin((xx + yy) / (1.7 * period) + ph[2]))
    ref = np.clip(np.rint(128 + 40 * t + rng.normal(0, 1.0, t.shape)), 0, 255).astype(np.uint8)
    return (255 - ref).astype(np.uint8), ref


# ---- geometry / work counts (prediction_frame.c:9-23, main.c:53-54,73-76) -----------------

def block_grid(width: int, height: int, blk_dim: int):
    """(x0, y0, w, h) int arrays for the raster block grid."""
    nbx, nby = -(-width // blk_dim), -(-height // blk_dim)
    bx = np.tile(np.arange(nbx), nby)
    by = np.repeat(np.arange(nby), nbx)
    x0, y0 = bx * blk_dim, by * blk_dim
    return x0, y0, np.minimum(blk_dim, width - x0), np.minimum(blk_dim, height - y0)


def _axis(n: int, b: int, r: int):
    p = np.arange(0, n, b)
    e = np.minimum(b, n - p)
    lo = np.maximum(0, p - r)
    hi = np.minimum(n - 1, p + e - 1 + r)
    return e, hi - e + 1 - lo + 1


def candidates(width: int, height: int, blk_dim: int, extra_span: int) -> int:
    _, cx = _axis(width, blk_dim, extra_span)
    _, cy = _axis(height, blk_dim, extra_span)
    return int(cx.sum()) * int(cy.sum())


def pixel_compares(width: int, height: int, blk_dim: int, extra_span: int) -> int:
    """Exact number of (cur-ref)^2 terms of one frame pair."""
    ex, cx = _axis(width, blk_dim, extra_span)
    ey, cy = _axis(height, blk_dim, extra_span)
    return int((ex * cx).sum()) * int((ey * cy).sum())

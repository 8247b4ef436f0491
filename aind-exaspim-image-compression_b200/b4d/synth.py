"""Seeded synthetic ExaSPIM-like uint16 volumes.

Grounded in constants of the reference tree: pedestal 37 counts
(scripts/evaluate_bm4dnet.py:207), sigma 24 (scripts/precompute.py:284),
PSF of 2-3 voxels (machine_learning/metrics.py:76-77), bright structures
800-60 000 counts (tests/test_metrics.py:24-33).  Clean signal = sparse blurred
poly-line fibres + a few soma-like blobs on a pedestal; noise = additive white
Gaussian so that sigma is known exactly.

Large volumes tile one clean 128^3 field periodically and draw noise per tile
from SeedSequence([seed, tz, ty, tx]); any sub-block can be regenerated on any
rank without materialising the whole volume.
"""
import numpy as np

PEDESTAL = 37.0
TILE = 128


def _blur_axis(a, sigma, axis):
    radius = int(4.0 * sigma + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    a = np.moveaxis(a, axis, -1)
    pad = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(radius, radius)], mode="wrap")
    out = np.zeros_like(a)
    for i, w in enumerate(k):
        out += w * pad[..., i : i + a.shape[-1]]
    return np.moveaxis(out, -1, axis)


def clean_tile(seed, size=TILE):
    """One periodic clean tile (float32, pedestal included)."""
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), 0xC1EA]))
    vol = np.zeros((size, size, size), dtype=np.float32)
    n_fibres = int(rng.integers(6, 13))
    for _ in range(n_fibres):
        amp = float(np.exp(rng.uniform(np.log(200.0), np.log(8000.0))))
        radius = int(rng.integers(1, 3))
        p = rng.uniform(0, size, size=3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        n_steps = int(rng.integers(size // 2, 2 * size))
        for _s in range(n_steps):
            d = d + 0.15 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            p = (p + d) % size
            c = np.floor(p).astype(int)
            for dz in range(-radius + 1, radius):
                for dy in range(-radius + 1, radius):
                    for dx in range(-radius + 1, radius):
                        vol[(c[0] + dz) % size, (c[1] + dy) % size, (c[2] + dx) % size] = amp
    for _ in range(int(rng.integers(0, 3))):
        amp = float(rng.uniform(3000.0, 30000.0))
        rad = float(rng.uniform(4.0, 8.0))
        c = rng.uniform(0, size, size=3)
        g = np.arange(size, dtype=np.float32)
        dz = np.minimum(np.abs(g - c[0]), size - np.abs(g - c[0]))[:, None, None]
        dy = np.minimum(np.abs(g - c[1]), size - np.abs(g - c[1]))[None, :, None]
        dx = np.minimum(np.abs(g - c[2]), size - np.abs(g - c[2]))[None, None, :]
        vol[(dz * dz + dy * dy + dx * dx) <= rad * rad] = amp
    v = vol.astype(np.float64)
    for axis, s in enumerate((1.5, 1.2, 1.2)):
        v = _blur_axis(v, s, axis)
    return (v + PEDESTAL).astype(np.float32)


def _noise_tile(seed, tz, ty, tx, sigma, size):
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), int(tz), int(ty), int(tx)]))
    return rng.standard_normal((size, size, size), dtype=np.float32) * np.float32(sigma)


def vol(D, H, W, seed, sigma=24.0, z0=0, y0=0, x0=0, clean=None):
    """uint16 (D,H,W) block of the seeded volume whose origin is (z0,y0,x0)."""
    if clean is None:
        clean = clean_tile(seed)
    T = clean.shape[0]
    out = np.empty((D, H, W), dtype=np.uint16)
    for tz in range(z0 // T, (z0 + D - 1) // T + 1):
        za, zb = max(z0, tz * T), min(z0 + D, (tz + 1) * T)
        for ty in range(y0 // T, (y0 + H - 1) // T + 1):
            ya, yb = max(y0, ty * T), min(y0 + H, (ty + 1) * T)
            for tx in range(x0 // T, (x0 + W - 1) // T + 1):
                xa, xb = max(x0, tx * T), min(x0 + W, (tx + 1) * T)
                sl = (slice(za - tz * T, zb - tz * T), slice(ya - ty * T, yb - ty * T), slice(xa - tx * T, xb - tx * T))
                noisy = clean[sl] + _noise_tile(seed, tz, ty, tx, sigma, T)[sl]
                out[za - z0 : zb - z0, ya - y0 : yb - y0, xa - x0 : xb - x0] = np.clip(
                    np.rint(noisy), 0, 65535
                ).astype(np.uint16)
    return out


def clean_vol(D, H, W, seed, z0=0, y0=0, x0=0):
    """The noise-free float32 field matching vol(...) (for PSNR-style checks)."""
    clean = clean_tile(seed)
    T = clean.shape[0]
    zi = (np.arange(z0, z0 + D) % T)[:, None, None]
    yi = (np.arange(y0, y0 + H) % T)[None, :, None]
    xi = (np.arange(x0, x0 + W) % T)[None, None, :]
    return clean[zi, yi, xi]

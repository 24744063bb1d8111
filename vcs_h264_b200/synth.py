"""Deterministic synthetic clips for tests and bench.py (recipe of SURVEY 8d).

A blurred random texture panned globally by (round(11 sin(2 pi t/16)), round(7 cos(2 pi t/12)))
plus i.i.d. integer noise in [-2,2]: every macroblock fails the static test (threshold 2000) and
the true displacement relative to the GOP's I-frame stays inside a +/-16 window.
"""
from __future__ import annotations

import numpy as np


def _blur_axis(a, sigma, axis):
    r = int(4 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    a = np.moveaxis(a, axis, 0)
    pad = np.concatenate([a[r:0:-1], a, a[-2:-r - 2:-1]], 0)
    out = np.zeros_like(a)
    for i, w in enumerate(k):
        out += w * pad[i:i + a.shape[0]]
    return np.moveaxis(out, 0, axis)


def texture(H, W, seed=1234, sigma=2.0, margin=96):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H + 2 * margin, W + 2 * margin, 3)).astype(np.float32)
    base = _blur_axis(_blur_axis(base, sigma, 0), sigma, 1)
    lo, hi = base.min(), base.max()
    return np.clip(np.rint((base - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def pan(t):
    return int(round(11 * np.sin(2 * np.pi * t / 16))), int(round(7 * np.cos(2 * np.pi * t / 12)))


def clip(T, H, W, seed=1234, noise=2, margin=96, out=None):
    """uint8 [T,H,W,3] BGR-interleaved, C-contiguous."""
    base = texture(H, W, seed, margin=margin)
    rng = np.random.default_rng(seed + 1)
    frames = out if out is not None else np.empty((T, H, W, 3), np.uint8)
    for t in range(T):
        dx, dy = pan(t)
        f = base[margin + dy:margin + dy + H, margin + dx:margin + dx + W].astype(np.int16)
        if noise:
            f = f + rng.integers(-noise, noise + 1, f.shape, dtype=np.int16)
        frames[t] = np.clip(f, 0, 255).astype(np.uint8)
    return frames


def still(H, W, seed=4321):
    """Synthetic still of a given shape (BASELINE config 4: bigImg.png is missing from the repo)."""
    return np.ascontiguousarray(texture(H, W, seed, margin=0))

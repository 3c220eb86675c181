"""Seeded synthetic 8-bit BGR sequences (the workload of SURVEY.md 8d).

Deterministic on the CPU (numpy only), frame-addressable so a frame-range
shard can build exactly its own frames: a band-limited textured background
under a piecewise-constant global pan, a handful of textured rectangles with
independent integer velocities, +-2 uniform pixel noise per frame, and one
exactly constant patch (together with the encoder's zero padding rows it
exercises the tie-break quirks of the reference search).
"""
from __future__ import annotations

import numpy as np


def _box_blur_wrap(a: np.ndarray, k: int, axis: int) -> np.ndarray:
    """Periodic box filter of odd width k along `axis` (float32)."""
    r = k // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r + 1, r)
    c = np.cumsum(np.pad(a, pad, mode="wrap"), axis=axis, dtype=np.float64)
    n = a.shape[axis]
    hi = np.take(c, np.arange(k, k + n), axis=axis)
    lo = np.take(c, np.arange(0, n), axis=axis)
    return ((hi - lo) / k).astype(np.float32)


def _texture(rng: np.random.Generator, h: int, w: int, k: int = 5) -> np.ndarray:
    """(h, w, 3) uint8 blurred noise, values kept inside [4, 251]."""
    base = rng.random((h, w), dtype=np.float32)
    own = rng.random((h, w, 3), dtype=np.float32)
    f = 0.7 * base[..., None] + 0.3 * own
    for _ in range(3):
        f = _box_blur_wrap(f, k, 0)
        f = _box_blur_wrap(f, k, 1)
    lo, hi = f.min(), f.max()
    f = (f - lo) / max(hi - lo, 1e-6)
    return (4.0 + f * 247.0 + 0.5).astype(np.uint8)


class SyntheticSequence:
    """`n_frames` BGR frames of `w` x `h`; `frame(i)` is independent of order."""

    def __init__(self, w: int, h: int, n_frames: int, seed: int = 1234,
                 n_rects: int = 6, max_pan: int = 6, noise: int = 2):
        self.w, self.h, self.n_frames = w, h, n_frames
        rng = np.random.default_rng(seed)
        self.bg = _texture(rng, h, w)
        # piecewise-constant pan velocity, changes every 16 frames
        seg = (n_frames + 15) // 16
        vel = rng.integers(-max_pan, max_pan + 1, size=(seg, 2))
        vel = np.repeat(vel, 16, axis=0)[:n_frames]
        vel[0] = 0
        self.pan = np.cumsum(vel, axis=0)
        # rectangles: own texture, size, start position, velocity
        self.rects = []
        for _ in range(n_rects):
            rh = int(rng.integers(max(8, h // 12), max(9, h // 4)))
            rw = int(rng.integers(max(8, w // 16), max(9, w // 5)))
            tex = _texture(rng, rh, rw, 3)
            pos = rng.integers(0, [max(1, w - rw), max(1, h - rh)])
            v = rng.integers(-8, 9, size=2)
            self.rects.append((tex, pos.astype(np.int64), v.astype(np.int64)))
        # noise tile in uint8 modular arithmetic (values -noise..noise)
        self.noise = rng.integers(-noise, noise + 1, size=(h, w, 3)).astype(np.uint8)
        self.noise_shift = rng.integers(0, [h, w], size=(n_frames, 2))
        # constant patch (>= 64x64 when the frame allows it)
        ps = min(96, h // 2, w // 2)
        self.flat = (h // 8, w // 8, ps, (40, 120, 200))

    def frame(self, i: int) -> np.ndarray:
        if not 0 <= i < self.n_frames:
            raise IndexError(i)
        w, h = self.w, self.h
        dx, dy = int(self.pan[i, 0]), int(self.pan[i, 1])
        f = np.roll(self.bg, (dy, dx), axis=(0, 1))
        for tex, pos, v in self.rects:
            rh, rw, _ = tex.shape
            x = int((pos[0] + v[0] * i) % max(1, w - rw))
            y = int((pos[1] + v[1] * i) % max(1, h - rh))
            f[y:y + rh, x:x + rw] = tex
        sy, sx = int(self.noise_shift[i, 0]), int(self.noise_shift[i, 1])
        f = f + np.roll(self.noise, (sy, sx), axis=(0, 1))  # uint8 wrap == signed add
        py, px, ps, col = self.flat
        f[py:py + ps, px:px + ps] = col
        return f

    def frames(self, lo: int = 0, hi: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        hi = self.n_frames if hi is None else hi
        if out is None:
            out = np.empty((hi - lo, self.h, self.w, 3), np.uint8)
        for k, i in enumerate(range(lo, hi)):
            out[k] = self.frame(i)
        return out

"""Seeded synthetic stereo pairs with known shifts.

``texture``  "noise": i.i.d. uniform uint8 noise, one uniform shift (adaptive P2 mostly collapses to P1).
             "scene": 3x3 box-blurred noise, one constant-intensity block and three horizontal bands
                      with shifts {D/8, D/3, 3D/4}: exercises adaptive P2 = P2_init, the uint8 wrap of
                      L_r, uniqueness failures and occlusion / LR-check failures.
Seeds follow the survey's convention ``0xB200 + frame index``.  Pure numpy; deterministic for a given
numpy major version (PCG64).
"""
from __future__ import annotations

import numpy as np


def _box3(a: np.ndarray) -> np.ndarray:
    p = np.pad(a.astype(np.uint32), 1, mode="edge")
    h, w = a.shape
    s = sum(p[i:i + h, j:j + w] for i in range(3) for j in range(3))
    return (s // 9).astype(np.uint8)


def make_pair(width: int, height: int, disp_range: int, seed: int = 0xB200, texture: str = "noise",
              shift: int | None = None) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Return (left, right, true_disparity) as uint8[H,W], uint8[H,W], int32[H,W].

    The right image is the same texture sampled ``shift`` columns further along, so the left pixel at
    column x matches the right pixel at column x - shift (disparity = shift)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w, h, d = width, height, disp_range
    if shift is None:
        shift = max(1, min(d - 2, (37 * d) // 128))
    tex = rng.integers(0, 256, size=(h, w + 2 * d), dtype=np.uint8)
    truth = np.full((h, w), shift, np.int32)
    if texture == "noise":
        left = tex[:, d:d + w].copy()
        right = tex[:, d + shift:d + shift + w].copy()
        return left, right, truth
    if texture != "scene":
        raise ValueError(texture)
    tex = _box3(tex)
    left = tex[:, d:d + w].copy()
    right = np.empty_like(left)
    bands = [max(1, d // 8), max(1, d // 3), max(1, (3 * d) // 4)]
    edges = [0, h // 3, (2 * h) // 3, h]
    for b, s in enumerate(bands):
        s = min(s, d - 2) if d > 2 else 0
        r0, r1 = edges[b], edges[b + 1]
        right[r0:r1] = tex[r0:r1, d + s:d + s + w]
        truth[r0:r1] = s
    bs = min(64, h // 2, w // 2)
    y0, x0 = h // 4, w // 2
    left[y0:y0 + bs, x0:x0 + bs] = 128
    right[y0:y0 + bs, max(0, x0 - bands[0]):max(0, x0 - bands[0]) + bs] = 128
    return left, right, truth

"""Independent numpy restatement of the cost and aggregation steps.  TEST INFRASTRUCTURE ONLY (see oracle/sgm_oracle.h).

A second opinion next to the C restatement (oracle/sgm_oracle.c): written from the reference's text, not from the C
port, vectorised over all paths of a direction, so a transcription error in one of the two shows up as a difference.
It is the only independent check of the 9x7 / 64-bit census extension's cost volume and aggregation (the extension has
no reference code: "parity unpinned").  Small shapes only (pure numpy, O(H*W*D) per direction).

Follows SemiGlobalMatching.c:134-159 (census; generalised to a CW x CH window with the same conventions),
:161-196 (cost), :229-372 (CostAggregate; path topology per SURVEY.md section 8a), :198-221 (sum over directions)."""
from __future__ import annotations

import numpy as np

DIRECTIONS = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1)]      # SGM.c:213-220


def census(img: np.ndarray, cw: int = 5, ch: int = 5) -> np.ndarray:
    """Bit = neighbour < centre, rows outer / columns inner, shifted in from the LSB; border stays 0 (SGM.c:134-159)."""
    h, w = img.shape
    out = np.zeros((h, w), np.uint64)
    rx, ry = cw // 2, ch // 2
    if w <= cw or h <= ch:
        return out
    a = img.astype(np.int32)
    core = a[ry:h - ry, rx:w - rx]
    acc = np.zeros(core.shape, np.uint64)
    for r in range(-ry, ry + 1):
        for c in range(-rx, rx + 1):
            acc = (acc << np.uint64(1)) | (a[ry + r:h - ry + r, rx + c:w - rx + c] < core).astype(np.uint64)
    out[ry:h - ry, rx:w - rx] = acc
    return out


def _popcount64(x: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(x).view(np.uint8).reshape(x.shape + (8,)), axis=-1).sum(axis=-1).astype(np.uint8)


def cost_volume(cl: np.ndarray, cr: np.ndarray, dmin: int, dmax: int) -> np.ndarray:
    """C(i,j,d) = popcount(cl[i,j] ^ cr[i,j-d]); 127 where j-d < 0 (SGM.c:161-183)."""
    h, w = cl.shape
    out = np.full((h, w, dmax - dmin), 127, np.uint8)
    for k, d in enumerate(range(dmin, dmax)):
        if d < w:
            out[:, d:, k] = _popcount64(cl[:, d:] ^ cr[:, :w - d])
    return out


def walk(w: int, h: int, dx: int, dy: int, i: int) -> list[int]:
    """Pixel indices of path i of direction (dx,dy), incl. out-of-image ones (SGM.c:232-367).
    `row`, `col` are the reference's uint16 trackers, which drift from the true position after a wrap."""
    fwd = (dx, dy) in ((1, 0), (0, 1), (1, 1), (-1, 1))
    dr = 1 if fwd else -1
    if dy == 0:
        pos = i * w if fwd else i * w + w - 1
    else:
        pos = i if fwd else (h - 1) * w + i
    out = [pos]
    row, col = (0 if fwd else h - 1), i
    for _ in range((w if dy == 0 else h) - 1):
        if dy == 0:
            pos += dr
        elif dx == 0:
            pos += dr * w
        elif (fwd and col == w - 1 and row < h - 1) or (not fwd and col == w - 1 and row > 0):
            pos = (row + dr) * w; col = 0
        elif (not fwd and col == 0 and row > 0) or (fwd and col == 0 and row < h - 1):
            pos = (row + dr) * w + w - 1; col = w - 1
        elif (dx, dy) in ((1, 1), (-1, -1)):
            pos += dr * (w + 1)
        else:
            pos += dr * (w - 1)
        out.append(pos)
        row = (row + dr) & 0xFFFF
        col = ((col - dr) if (dx, dy) in ((-1, 1), (1, -1)) else (col + dr)) & 0xFFFF
    return out


def aggregate_direction(img: np.ndarray, cost: np.ndarray, dx: int, dy: int, p1: int, p2_init: int) -> np.ndarray:
    """Contribution of one direction to S as uint16 [H,W,D] (SGM.c:229-372).  All paths advance in lock step; a pixel
    visited twice (irregular diagonal paths) accumulates both visits; out-of-image visits are skipped (they are always
    the last of their path, so nothing depends on them)."""
    h, w, dd = cost.shape
    n = h * w
    npaths = h if dy == 0 else w
    paths = np.array([walk(w, h, dx, dy, i) for i in range(npaths)], np.int64)       # [npaths, steps]
    g = img.reshape(-1).astype(np.int32)
    c = cost.reshape(n, dd).astype(np.int32)
    out = np.zeros((n, dd), np.int64)
    pos = paths[:, 0]
    lp = c[pos].copy()                                                               # first pixel: L = C (:266-275)
    np.add.at(out, pos, lp)
    gprev = g[pos]
    alive = np.ones(npaths, bool)
    for s in range(1, paths.shape[1]):
        pos = paths[:, s]
        ok = alive & (pos >= 0) & (pos < n)
        alive = ok
        if not ok.any():
            break
        q = np.where(ok, pos, 0)
        mn = lp.min(axis=1, keepdims=True)
        pad = np.full((npaths, 1), 255, np.int32)                                   # Lp[-1] = Lp[D] = 255 (:260-263)
        l2 = np.concatenate([pad, lp[:, :-1]], axis=1) + p1
        l3 = np.concatenate([lp[:, 1:], pad], axis=1) + p1
        p2 = np.maximum(p1, p2_init // (np.abs(g[q] - gprev) + 1))[:, None]         # :335
        m = np.minimum(np.minimum(lp, l2), np.minimum(l3, mn + p2))
        new = (c[q] + m - mn) & 0xFF                                                # (uint8_t) truncation (:343)
        lp = np.where(ok[:, None], new, lp)
        gprev = np.where(ok, g[q], gprev)
        np.add.at(out, q[ok], new[ok])
    return out.reshape(h, w, dd).astype(np.uint16)


def aggregate(img: np.ndarray, cost: np.ndarray, p1: int, p2_init: int, num_paths: int = 8):
    dirs = DIRECTIONS[:4] if num_paths == 4 else DIRECTIONS
    per = [aggregate_direction(img, cost, dx, dy, p1, p2_init) for dx, dy in dirs]
    return sum(a.astype(np.uint32) for a in per).astype(np.uint16), per

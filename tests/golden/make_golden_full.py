#!/usr/bin/env python3
"""Generate the FULL-SIZE golden fixtures (BASELINE.json configs C3, C4, C5) from the REFERENCE ITSELF.

Run in the build container (needs /root/reference, ~40 GB of RAM, ~15 min on 8 cores):

    python tests/golden/make_golden_full.py [c3] [c5] [c4]

* ``full_c3.npz``  2864x1924, D=256, 8 paths -- the reference's SemiGlobalMatching.c compiled verbatim by
  ``full_c5.npz``  3840x2160, D=256, 8 paths    oracle/build_ref.py (d256 one-token patch of SGM.c:272, guard rows,
                   -mcmodel=large), run through its own stage functions on `make_pair(seed=0xB200)` for both
                   textures.  The volumes are far too large to commit (S is 2.8 / 4.2 GB), so per texture the
                   fixture holds: md5 of both input images (guards generator drift), md5 + sum + per-row sums of S,
                   and for every disparity stage md5, number of valid pixels, float64 sum of the valid pixels,
                   CRC32 of every row and of every column (a mismatch is localised to (row, column) candidates
                   without the stored map), plus every 32nd row of the post-LR and final maps in full.
* ``full_c4.npz``  all 256 pairs of config C4 (1242x375, D=128, 4 paths; p4 patch of SGM.c:217-220): md5 of the
                   hot-path result (post-LR disparity) of every pair + md5 of its inputs.  Pair k uses
                   seed 0xB200 + k, texture "scene" when k % 3 == 0 else "noise".
"""
from __future__ import annotations

import ctypes as C
import hashlib
import multiprocessing as mp
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pyoracle import Reference, _padded, _ptr, options  # noqa: E402
from soc_project_stereo_matching_b200.synth import make_pair  # noqa: E402

FULL = {"c3": (2864, 1924, 256), "c5": (3840, 2160, 256)}
C4 = (1242, 375, 128, 256)
DISP_STAGES = ["disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final"]
ROW_STEP = 32


def md5(a: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def row_crc(a: np.ndarray) -> np.ndarray:
    return np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in a], np.uint32)


def disp_summary(prefix: str, a: np.ndarray) -> dict:
    v = np.isfinite(a)
    return {f"{prefix}_md5": np.asarray(md5(a)), f"{prefix}_valid": np.asarray(int(v.sum())),
            f"{prefix}_sum": np.asarray(float(a[v].astype(np.float64).sum())),
            f"{prefix}_rowcrc": row_crc(a), f"{prefix}_colcrc": row_crc(np.ascontiguousarray(a.T))}


def c4_texture(k: int) -> str:
    return "scene" if k % 3 == 0 else "noise"


def run_full(job):
    name, tex = job
    w, h, d = FULL[name]
    o = options(max_disparity=d)
    t0 = time.time()
    left, right, _ = make_pair(w, h, d, seed=0xB200, texture=tex)
    ref = Reference(w, h, d)
    ref._init(o)
    lb, lv = _padded(left); rb, rv = _padded(right)
    aggr = np.empty((h, w, d), np.uint16)
    outs = {k: np.empty((h, w), np.float32) for k in DISP_STAGES}
    # census_l, census_r, cost (not copied out: 1.4 / 2.1 GB, recomputable), aggr, then the five disparity stages
    args = [_ptr(lv), _ptr(rv), None, None, None, _ptr(aggr)] + [_ptr(outs[k]) for k in DISP_STAGES]
    ok = ref._call_big_stack(lambda: ref.lib.ref_match_staged(*args))
    assert ok
    res = {f"{tex}_md5_left": np.asarray(md5(left)), f"{tex}_md5_right": np.asarray(md5(right)),
           f"{tex}_aggr_md5": np.asarray(md5(aggr)), f"{tex}_aggr_sum": np.asarray(int(aggr.sum(dtype=np.uint64))),
           f"{tex}_aggr_max": np.asarray(int(aggr.max())),
           f"{tex}_aggr_rowsum": aggr.reshape(h, -1).sum(axis=1, dtype=np.uint64)}
    for k in DISP_STAGES:
        res.update(disp_summary(f"{tex}_{k}", outs[k]))
    for k in ("disp_lr", "disp_final"):
        res[f"{tex}_{k}_rows"] = outs[k][::ROW_STEP].copy()
    print(f"{name}/{tex}: {time.time() - t0:.0f} s, valid {float(np.isfinite(outs['disp_lr']).mean()):.4f}", flush=True)
    return name, res


def run_c4(k):
    w, h, d, _ = C4
    o = options(max_disparity=d, num_paths=4)
    left, right, _ = make_pair(w, h, d, seed=0xB200 + k, texture=c4_texture(k))
    ref = run_c4.ref if hasattr(run_c4, "ref") else Reference(w, h, d, "p4")
    run_c4.ref = ref
    out = ref.hotpath(left, right, o)
    return k, md5(out), md5(left), md5(right), int(np.isfinite(out).sum())


def main(argv) -> None:
    which = [a for a in argv if a in ("c3", "c5", "c4")] or ["c3", "c5", "c4"]
    ctx = mp.get_context("fork")
    jobs = [(n, t) for n in which if n in FULL for t in ("noise", "scene")]
    if jobs:
        for n in {j[0] for j in jobs}:
            Reference(*FULL[n])                       # compile once, before the workers fork
        with ctx.Pool(len(jobs)) as pool:
            merged: dict = {}
            for name, res in pool.imap_unordered(run_full, jobs):
                merged.setdefault(name, {}).update(res)
        for name, res in merged.items():
            w, h, d = FULL[name]
            np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), shape=np.asarray([w, h, d]), row_step=np.asarray(ROW_STEP), **res)
            print("wrote", f"full_{name}.npz")
    if "c4" in which:
        w, h, d, n = C4
        Reference(w, h, d, "p4")
        t0 = time.time()
        with ctx.Pool(os.cpu_count() or 1) as pool:
            rows = sorted(pool.map(run_c4, range(n), chunksize=4))
        np.savez_compressed(os.path.join(HERE, "full_c4.npz"), shape=np.asarray([w, h, d, n]),
                            md5_hotpath=np.asarray([r[1] for r in rows]), md5_left=np.asarray([r[2] for r in rows]),
                            md5_right=np.asarray([r[3] for r in rows]), valid=np.asarray([r[4] for r in rows]))
        print(f"wrote full_c4.npz ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main(sys.argv[1:])

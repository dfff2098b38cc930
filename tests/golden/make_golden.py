#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

For every case the reference's own SemiGlobalMatching.c -- compiled verbatim by oracle/build_ref.py
("sanitised oracle": guard rows + padded inputs; d256 / p4 one-token patches where the case needs
them) -- is run and every stage is recorded:

* ``cone.npz``      config C1: the bundled Middlebury cone pair (Data/cone/im2.png, im6.png) converted
                    to grey exactly like main.c does (stb_image: (77R+150G+29B)>>8), main.c's options.
                    Holds the two grey images, the post-LR and final disparities and the md5 of every
                    stage (the raw S volume would be 21 MB).
* ``small_*.npz``   seeded synthetic pairs, small enough to store every stage in full (md5 only for
                    the cost volume and the per-direction path costs), covering D=256
                    (d256 patch), 4 paths (p4 patch), min_disparity > 0, flags off, odd sizes.
* ``cone_demo.npz`` the 8-bit image the reference demo wrote (Data/cone/im2.d.png) -- the only
                    known-answer artefact in the reference tree.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pyoracle import Reference, options, stb_gray  # noqa: E402
from soc_project_stereo_matching_b200.synth import make_pair  # noqa: E402

DATA = "/root/reference/SemiGlobalMatching/Data"

# name, W, H, texture, seed, option overrides
SMALL_CASES = [
    ("small_d16", 48, 32, "scene", 0xB200, dict(max_disparity=16)),
    ("small_d64_noise", 64, 40, "noise", 0xB201, dict(max_disparity=64)),
    ("small_d128", 96, 24, "scene", 0xB202, dict(max_disparity=128)),
    ("small_d256", 80, 20, "scene", 0xB203, dict(max_disparity=256)),
    ("small_p4", 50, 30, "scene", 0xB204, dict(max_disparity=32, num_paths=4)),
    ("small_mind", 64, 36, "scene", 0xB205, dict(min_disparity=5, max_disparity=45)),
    ("small_flags_off", 40, 40, "noise", 0xB206, dict(max_disparity=24, check_unique=False, check_lr=False, remove_speckles=False)),
    ("small_odd", 37, 23, "scene", 0xB207, dict(max_disparity=21, p1=7, p2_init=90, uniqueness_ratio=0.95, lrcheck_thres=0.5, min_speckle_area=12)),
    ("small_square", 32, 32, "scene", 0xB208, dict(max_disparity=20)),
    ("small_tiny", 7, 6, "noise", 0xB209, dict(max_disparity=4)),
]

STAGES = ["census_left", "census_right", "cost", "aggr", "disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final"]


def md5(a: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def opts_to_arrays(o: dict) -> dict:
    return {f"opt_{k}": np.asarray(v) for k, v in o.items()}


def main() -> None:
    import cv2

    # ---- C1: cone
    left = stb_gray(cv2.imread(f"{DATA}/cone/im2.png")[..., ::-1])
    right = stb_gray(cv2.imread(f"{DATA}/cone/im6.png")[..., ::-1])
    o = options()
    h, w = left.shape
    ref = Reference(w, h, o["max_disparity"] - o["min_disparity"])
    res = ref.match(left, right, o, per_direction=True)
    plain = ref.match_plain(left, right, o)
    assert np.array_equal(plain.view(np.uint32), res["disp_final"].view(np.uint32))
    out = dict(left=left, right=right, disp_lr=res["disp_lr"], disp_final=res["disp_final"], **opts_to_arrays(o))
    for k in STAGES:
        out[f"md5_{k}"] = np.asarray(md5(res[k]))
    for i, a in enumerate(res["path_cost"]):
        out[f"md5_path_cost_{i}"] = np.asarray(md5(a))
    out["aggr_sum"] = np.asarray(res["aggr"].sum(dtype=np.uint64))
    out["aggr_max"] = np.asarray(res["aggr"].max())
    np.savez_compressed(os.path.join(HERE, "cone.npz"), **out)
    print("cone:", {k: str(v) for k, v in out.items() if k.startswith("md5_")})

    demo = cv2.imread(f"{DATA}/cone/im2.d.png", 0)
    np.savez_compressed(os.path.join(HERE, "cone_demo.npz"), demo=demo)

    # ---- small synthetic cases, all stages in full
    for name, w, h, tex, seed, kw in SMALL_CASES:
        o = options(**kw)
        d = o["max_disparity"] - o["min_disparity"]
        left, right, _ = make_pair(w, h, d, seed=seed, texture=tex)
        ref = Reference(w, h, d, "p4" if o["num_paths"] == 4 else "")
        res = ref.match(left, right, o, per_direction=True)
        out = dict(left=left, right=right, **opts_to_arrays(o))
        for k in STAGES:
            if k in res:
                if k == "cost":
                    out["md5_cost"] = np.asarray(md5(res[k]))      # recomputable from the census taps
                else:
                    out[k] = res[k]
        for i, a in enumerate(res["path_cost"]):
            out[f"md5_path_cost_{i}"] = np.asarray(md5(a))         # their sum is the stored `aggr`
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "valid fraction", float(np.isfinite(res["disp_final"]).mean()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate tests/golden/eval_depth.npz from the reference's own Python evaluation code
(/root/reference/HostScript_Server/depth_image.py, imported unmodified; runs only where /root/reference exists).

Vectors for the "next" row N4 of SURVEY.md section 8f:
  * disparity_to_depth (depth_image.py:138-165) on seeded float32 disparity maps with NaN holes;
  * compare_img (depth_image.py:276-319): (rmse, bpr, n_valid) for several (ground truth, test, threshold) triples.

    python tests/golden/make_golden_eval.py
"""
import logging
import os
import sys
import types

import numpy as np

REF = "/root/reference/HostScript_Server"
sys.path.insert(0, REF)
import depth_image  # noqa: E402  (the reference module)

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> int:
    rng = np.random.Generator(np.random.PCG64(0xE7A1))
    log = logging.getLogger("golden")
    out = {}
    cases = []
    for k, (h, w, baseline, fx, doffs) in enumerate([(48, 64, 193.001, 3979.911, 124.343), (33, 130, 111.53, 1758.23, 0.0),
                                                     (180, 320, 536.62, 7190.247 * 0.1, 342.789 * 0.1)]):
        disp = (rng.random((h, w), dtype=np.float32) * np.float32(250.0) + np.float32(0.5)).astype(np.float32)
        disp[rng.random((h, w)) < 0.15] = np.nan
        cam0 = np.array([[fx, 0, 100.0], [0, fx, 50.0], [0, 0, 1]], dtype=np.float32)     # parse_3x3_float_matrix returns float32
        calib = types.SimpleNamespace(cam0=cam0, cam1=cam0.copy(), baseline=float(baseline), doffs=float(doffs))
        depth = depth_image.disparity_to_depth(disp, calib, 0)
        assert depth.dtype == np.float32
        out[f"d2d{k}_disp"] = disp
        out[f"d2d{k}_params"] = np.array([np.float32(baseline), cam0[0, 0], np.float32(doffs)], np.float32)
        out[f"d2d{k}_depth"] = depth
        # a noisy, partly invalid test map against this depth as ground truth
        test = depth * (np.float32(1.0) + rng.standard_normal((h, w)).astype(np.float32) * np.float32(0.01))
        test[rng.random((h, w)) < 0.1] = np.inf
        test[rng.random((h, w)) < 0.05] = np.nan
        for thr in (10.0, 100.0):
            rmse, bpr, nv = depth_image.compare_img(depth, test, log, abs_thresh=thr)
            cases.append((k, thr, rmse, bpr, nv))
        out[f"cmp{k}_test"] = test.astype(np.float32)
    # degenerate: no valid pixel
    gt = np.full((4, 5), np.nan, np.float32)
    rmse, bpr, nv = depth_image.compare_img(gt, np.ones((4, 5), np.float32), log)
    assert np.isnan(rmse) and np.isnan(bpr) and nv == 0
    out["cmp_cases"] = np.array(cases, np.float64)      # rows: case index, threshold, rmse, bpr, n_valid
    np.savez_compressed(os.path.join(HERE, "eval_depth.npz"), **out)
    print("wrote eval_depth.npz:", {k: v.shape for k, v in out.items()})
    return 0


if __name__ == "__main__":
    sys.exit(main())

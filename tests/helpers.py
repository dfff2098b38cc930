"""Shared helpers of the parity tests."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLOAT_STAGES = ("disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final")
OPT_KEYS = ("num_paths", "min_disparity", "max_disparity", "check_unique", "uniqueness_ratio", "check_lr",
            "lrcheck_thres", "remove_speckles", "min_speckle_area", "p1", "p2_init", "median")


def md5(a: np.ndarray) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_names(prefix: str = "small_") -> list[str]:
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name: str):
    """-> (left, right, opts dict in pyoracle.options() form, dict of stored stages / md5 strings)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    opts = {}
    for k in OPT_KEYS:
        v = z["opt_" + k]
        opts[k] = v.item()
    stages = {k: (str(z[k]) if k.startswith("md5_") else z[k]) for k in z.files
              if not k.startswith("opt_") and k not in ("left", "right")}
    return z["left"], z["right"], opts, stages


def to_sgm_option(opts: dict):
    """pyoracle-style option dict -> the product's SGMOption (reference ABI)."""
    import soc_project_stereo_matching_b200 as sgm
    return sgm.SGMOption(num_paths=opts["num_paths"], min_disparity=opts["min_disparity"], max_disparity=opts["max_disparity"],
                         is_check_unique=bool(opts["check_unique"]), uniqueness_ratio=opts["uniqueness_ratio"],
                         is_check_lr=bool(opts["check_lr"]), lrcheck_thres=opts["lrcheck_thres"],
                         is_remove_speckles=bool(opts["remove_speckles"]), min_speckle_area=opts["min_speckle_area"],
                         p1=opts["p1"], p2_init=opts["p2_init"])


def assert_same(name: str, got: np.ndarray, want: np.ndarray) -> None:
    """Bit-exact comparison (floats compared by bit pattern, so +inf == +inf and -0 != 0)."""
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{name}: dtype {got.dtype} != {want.dtype}"
    g = got.view(np.uint32) if got.dtype == np.float32 else got
    w = want.view(np.uint32) if want.dtype == np.float32 else want
    if np.array_equal(g, w):
        return
    bad = np.argwhere(g != w)
    first = tuple(int(x) for x in bad[0])
    raise AssertionError(f"{name}: {len(bad)} of {g.size} elements differ; first at {first}: got {got[first]!r}, want {want[first]!r}; "
                         f"rows {sorted(set(int(b[0]) for b in bad))[:8]} cols {sorted(set(int(b[1]) for b in bad))[:8]}")


def gpu_available() -> bool:
    try:
        import soc_project_stereo_matching_b200 as sgm
        return sgm.lib.SGMB_DeviceCount() > 0
    except Exception:
        return False

"""ctypes front-ends for the two parity checkers.  TEST INFRASTRUCTURE ONLY (see oracle/sgm_oracle.h):
importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never from the
product package.

* :class:`Oracle`   -- oracle/sgm_oracle.c, our restatement ("port"); any shape, re-entrant.
* :class:`Reference` -- the reference's own SemiGlobalMatching.c, compiled verbatim per shape by
  oracle/build_ref.py ("sanitised oracle"); one global instance per loaded library, landscape
  shapes only (the reference has undefined behaviour for H > W).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_ref  # noqa: E402

DIRECTIONS = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1)]  # SGM.c:213-220

DEFAULTS = dict(num_paths=8, min_disparity=0, max_disparity=64, check_unique=True, uniqueness_ratio=0.99,
                check_lr=True, lrcheck_thres=1.0, remove_speckles=True, min_speckle_area=50, p1=10,
                p2_init=150, median=True,  # main.c:48-65
                census_w=5, census_h=5)    # 5x5 = SGM.c:134-159; 9x7 = 64-bit extension, parity unpinned


def options(**kw) -> dict:
    o = dict(DEFAULTS)
    unknown = set(kw) - set(o)
    if unknown:
        raise KeyError(unknown)
    o.update(kw)
    return o


class _Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("min_disparity", C.c_int32),
                ("max_disparity", C.c_int32), ("num_paths", C.c_int32), ("p1", C.c_int32),
                ("p2_init", C.c_int32), ("check_unique", C.c_int32), ("uniqueness_ratio", C.c_float),
                ("check_lr", C.c_int32), ("lrcheck_thres", C.c_float), ("remove_speckles", C.c_int32),
                ("min_speckle_area", C.c_int32), ("median", C.c_int32), ("census_w", C.c_int32),
                ("census_h", C.c_int32)]


class _Taps(C.Structure):
    _fields_ = [("census_left", C.c_void_p), ("census_right", C.c_void_p), ("cost", C.c_void_p),
                ("path_cost", C.c_void_p * 8), ("aggr", C.c_void_p), ("disp_left_wta", C.c_void_p),
                ("disp_right", C.c_void_p), ("disp_lr", C.c_void_p), ("disp_speckle", C.c_void_p),
                ("census64_left", C.c_void_p), ("census64_right", C.c_void_p)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Our C restatement.  ``match`` returns a dict of stages (numpy arrays)."""

    def __init__(self):
        self.lib = C.CDLL(build_ref.build_oracle())
        self.lib.sgmo_match.restype = C.c_int
        self.lib.sgmo_match.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_Taps)]
        self.lib.sgmo_match_hotpath.restype = C.c_int
        self.lib.sgmo_match_hotpath.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.sgmo_walk_path.restype = C.c_int

    @staticmethod
    def _params(w, h, o) -> _Params:
        return _Params(w, h, o["min_disparity"], o["max_disparity"], o["num_paths"], o["p1"], o["p2_init"],
                       int(o["check_unique"]), o["uniqueness_ratio"], int(o["check_lr"]), o["lrcheck_thres"],
                       int(o["remove_speckles"]), o["min_speckle_area"], int(o.get("median", True)),
                       int(o.get("census_w", 5)), int(o.get("census_h", 5)))

    def match(self, left: np.ndarray, right: np.ndarray, opts: dict, stages: bool = True,
              per_direction: bool = False) -> dict:
        h, w = left.shape
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        d = opts["max_disparity"] - opts["min_disparity"]
        prm = self._params(w, h, opts)
        out = {"disp_final": np.empty((h, w), np.float32)}
        taps = _Taps()
        if stages:
            out.update(census_left=np.zeros((h, w), np.uint32), census_right=np.zeros((h, w), np.uint32),
                       cost=np.empty((h, w, d), np.uint8), aggr=np.empty((h, w, d), np.uint16),
                       disp_left_wta=np.empty((h, w), np.float32), disp_lr=np.empty((h, w), np.float32),
                       disp_speckle=np.empty((h, w), np.float32))
            if opts["check_lr"]:
                out["disp_right"] = np.empty((h, w), np.float32)
            if (opts.get("census_w", 5), opts.get("census_h", 5)) != (5, 5):
                # 64-bit descriptors replace the uint32 census taps (extension, parity unpinned)
                out["census_left"] = np.zeros((h, w), np.uint64); out["census_right"] = np.zeros((h, w), np.uint64)
                taps.census64_left = _ptr(out["census_left"]); taps.census64_right = _ptr(out["census_right"])
            for k in ("census_left", "census_right", "cost", "aggr", "disp_left_wta", "disp_right", "disp_lr", "disp_speckle"):
                if k.startswith("census") and out[k].dtype == np.uint64:
                    continue
                setattr(taps, k, _ptr(out.get(k)))
            if per_direction:
                n = 4 if opts["num_paths"] == 4 else 8
                out["path_cost"] = [np.empty((h, w, d), np.uint16) for _ in range(n)]
                for i, a in enumerate(out["path_cost"]):
                    taps.path_cost[i] = a.ctypes.data
        rc = self.lib.sgmo_match(C.byref(prm), _ptr(left), _ptr(right), _ptr(out["disp_final"]), C.byref(taps))
        if rc != 0:
            raise ValueError(f"sgmo_match rejected the arguments (rc={rc})")
        return out

    def hotpath(self, left, right, opts) -> np.ndarray:
        h, w = left.shape
        prm = self._params(w, h, opts)
        out = np.empty((h, w), np.float32)
        rc = self.lib.sgmo_match_hotpath(C.byref(prm), _ptr(np.ascontiguousarray(left)), _ptr(np.ascontiguousarray(right)), _ptr(out))
        if rc != 0:
            raise ValueError(rc)
        return out

    def walk(self, w, h, dx, dy, path) -> np.ndarray:
        buf = np.empty(max(w, h), np.int64)
        n = self.lib.sgmo_walk_path(w, h, dx, dy, path, _ptr(buf))
        return buf[:n].copy()


class SGMOption(C.Structure):
    """Reference ABI, SemiGlobalMatching.h:24-40 (sizeof 28 on x86-64)."""
    _fields_ = [("num_paths", C.c_uint8), ("min_disparity", C.c_uint16), ("max_disparity", C.c_uint16),
                ("is_check_unique", C.c_bool), ("uniqueness_ratio", C.c_float), ("is_check_lr", C.c_bool),
                ("lrcheck_thres", C.c_float), ("is_remove_speckles", C.c_bool), ("min_speckle_area", C.c_uint16),
                ("p1", C.c_int16), ("p2_init", C.c_int16)]


def sgm_option(o: dict) -> SGMOption:
    return SGMOption(o["num_paths"], o["min_disparity"], o["max_disparity"], o["check_unique"], o["uniqueness_ratio"],
                     o["check_lr"], o["lrcheck_thres"], o["remove_speckles"], o["min_speckle_area"], o["p1"], o["p2_init"])


def _padded(img: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Embed an image in a buffer with one spare row before and after: the reference reads up to
    H-2 pixels outside the image on two diagonal paths."""
    h, w = img.shape
    pad = max(w, h)
    buf = np.zeros(pad + h * w + pad, np.uint8)
    view = buf[pad:pad + h * w].reshape(h, w)
    view[:] = img
    return buf, view


class Reference:
    """The reference C code itself, built for one (W, H, D[, p4]) shape.  ``available`` is False when
    neither /root/reference nor a prebuilt library exists (e.g. an unexpected shape on the GPU box)."""

    def __init__(self, w: int, h: int, d: int, variant: str = ""):
        if h > w:
            raise ValueError("reference has undefined behaviour for portrait images; use Oracle")
        self.w, self.h, self.d, self.variant = w, h, d, variant
        path = build_ref.build_ref(w, h, d, variant)
        self.available = path is not None and os.path.isfile(path)
        if not self.available:
            return
        self.lib = C.CDLL(path)
        assert self.lib.ref_sizeof_option() == C.sizeof(SGMOption) == 28
        assert (self.lib.ref_max_width(), self.lib.ref_max_disp()) == (w, d)
        self.lib.ref_gap_init_to_aggr.restype = C.c_long
        self.lib.ref_gap_left_to_init.restype = C.c_long
        # guard layout: cost_init directly precedes cost_aggr, census_left precedes cost_init
        rows = self.lib.ref_max_height()
        assert self.lib.ref_gap_init_to_aggr() >= w * rows * d > 0
        assert self.lib.ref_gap_left_to_init() >= w * rows * 4 > 0
        assert (rows - h) * w >= 2 * (h - 1)
        self.lib.SGM_Initialize.restype = C.c_bool
        self.lib.SGM_Initialize.argtypes = [C.c_uint16, C.c_uint16, C.POINTER(SGMOption)]
        self.lib.SGM_Match.restype = C.c_bool
        self.lib.SGM_Match.argtypes = [C.c_void_p] * 3
        self.lib.ref_match_staged.argtypes = [C.c_void_p] * 11
        self.lib.ref_match_hotpath.argtypes = [C.c_void_p] * 3
        self.lib.ref_aggregate_dir.argtypes = [C.c_int, C.c_int, C.c_void_p]

    def _call_big_stack(self, fn):
        """RemoveSpeckles keeps 5 bytes/pixel on the stack (SGM.c:588-589)."""
        if self.w * self.h * 5 < (4 << 20):
            return fn()
        res = []
        old = threading.stack_size(self.w * self.h * 6 + (16 << 20))
        try:
            t = threading.Thread(target=lambda: res.append(fn()))
            t.start(); t.join()
        finally:
            threading.stack_size(old)
        return res[0]

    def _init(self, opts):
        d = opts["max_disparity"] - opts["min_disparity"]
        assert d == self.d and (self.variant == "p4") == (opts["num_paths"] == 4)
        if not self.lib.SGM_Initialize(self.w, self.h, C.byref(sgm_option(opts))):
            raise ValueError("SGM_Initialize returned false")

    def match_plain(self, left, right, opts) -> np.ndarray:
        """Exactly main.c's use: SGM_Initialize + SGM_Match."""
        self._init(opts)
        lb, lv = _padded(left); rb, rv = _padded(right)
        out = np.empty((self.h, self.w), np.float32)
        ok = self._call_big_stack(lambda: self.lib.SGM_Match(_ptr(lv), _ptr(rv), _ptr(out)))
        assert ok
        return out

    def match(self, left, right, opts, per_direction: bool = False) -> dict:
        self._init(opts)
        h, w, d = self.h, self.w, self.d
        lb, lv = _padded(left); rb, rv = _padded(right)
        out = dict(census_left=np.empty((h, w), np.uint32), census_right=np.empty((h, w), np.uint32),
                   cost=np.empty((h, w, d), np.uint8), aggr=np.empty((h, w, d), np.uint16),
                   disp_left_wta=np.empty((h, w), np.float32), disp_right=np.empty((h, w), np.float32),
                   disp_lr=np.empty((h, w), np.float32), disp_speckle=np.empty((h, w), np.float32),
                   disp_final=np.empty((h, w), np.float32))
        order = ["census_left", "census_right", "cost", "aggr", "disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final"]
        ok = self._call_big_stack(lambda: self.lib.ref_match_staged(_ptr(lv), _ptr(rv), *[_ptr(out[k]) for k in order]))
        assert ok
        if not opts["check_lr"]:
            del out["disp_right"]
        if per_direction:
            n = 4 if opts["num_paths"] == 4 else 8
            out["path_cost"] = []
            for dx, dy in DIRECTIONS[:n]:
                a = np.empty((h, w, d), np.uint16)
                assert self.lib.ref_aggregate_dir(dx, dy, _ptr(a))
                out["path_cost"].append(a)
        return out

    def hotpath(self, left, right, opts) -> np.ndarray:
        self._init(opts)
        lb, lv = _padded(left); rb, rv = _padded(right)
        out = np.empty((self.h, self.w), np.float32)
        assert self.lib.ref_match_hotpath(_ptr(lv), _ptr(rv), _ptr(out))
        return out


def stb_gray(rgb: np.ndarray) -> np.ndarray:
    """stb_image's RGB -> 1-channel conversion used by main.c:25-26 (stb_image.h:1746-1749)."""
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((r * 77 + g * 150 + b * 29) >> 8).astype(np.uint8)


# ------------------------------------------------------------------ steps either side of the path
def board_gray(bgr_planes: np.ndarray) -> np.ndarray:
    """The board's colour -> grey conversion, ZedBoard/Vitis/lwip_tcp_perf_client/src/stereo_matching.c:18-24:
    (76*R + 150*G + 29*B) >> 8 on planar B,G,R input [3, H, W]."""
    b, g, r = (bgr_planes[i].astype(np.uint32) for i in range(3))
    return ((76 * r + 150 * g + 29 * b) >> 8).astype(np.uint8)


def stb_gray_planar(bgr_planes: np.ndarray) -> np.ndarray:
    """stb_image's weights (stb_image.h:1746-1749) on planar B,G,R input."""
    return stb_gray(np.stack([bgr_planes[2], bgr_planes[1], bgr_planes[0]], axis=-1))


def disparity_to_depth(disp: np.ndarray, baseline: float, fx: float, doffs: float) -> np.ndarray:
    """HostScript_Server/depth_image.py:138-165 restated: float32 throughout (cam matrices are float32 arrays,
    stereo_calibration.py:38; Python floats are weak scalars), depth = fl(fl(baseline*fx) / fl(disp + doffs)).
    Pinned by tests/golden/eval_depth.npz, generated from the reference module itself."""
    d = np.asarray(disp, np.float32)
    bf = np.float32(baseline) * np.float32(fx)
    return (bf / (d + np.float32(doffs))).astype(np.float32)


def compare_img(ground_truth: np.ndarray, test: np.ndarray, abs_thresh: float = 10.0):
    """HostScript_Server/depth_image.py:276-319 restated: (rmse, bpr, n_valid) over pixels finite in both maps."""
    valid = np.isfinite(test) & np.isfinite(ground_truth)
    n = int(np.count_nonzero(valid))
    if n == 0:
        return float("nan"), float("nan"), 0
    diff = test[valid] - ground_truth[valid]
    return float(np.sqrt(np.mean(np.square(diff)))), float(np.count_nonzero(np.abs(diff) > abs_thresh) / n), n

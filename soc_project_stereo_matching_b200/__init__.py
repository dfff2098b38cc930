"""B200-native Semi-Global Matching hot path behind the reference's C API.

The product is ``lib/libsgm_b200.so`` (C-ABI, hand-written sm_100a CUDA kernels; headers in
``include/``).  This package is the thin Python mirror of that ABI used by the tests and the
benchmark: same function names, argument meaning and error behaviour as the reference's
``SemiGlobalMatching.h`` (``SGM_Initialize`` / ``SGM_Reset`` / ``SGM_Match`` return ``False`` on
failure, never raise), plus the additive ``SGMB_*`` context API.

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every
compute entry point fails when no B200 is present, and loading fails loudly when the shared
library has not been built (``python -m soc_project_stereo_matching_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


__all__ = ["SGMOption", "SGM_Initialize", "SGM_Reset", "SGM_Match", "Context", "lib", "last_error",
           "default_option", "INVALID_FLOAT", "PIPE_SPECKLE", "PIPE_MEDIAN", "PIPE_TAPS", "PIPE_REFERENCE",
           "PIPE_HOTPATH", "STAGES"]

INVALID_FLOAT = np.float32(np.inf)                      # SemiGlobalMatching.h:12

PIPE_SPECKLE, PIPE_MEDIAN, PIPE_TAPS = 1, 2, 4
PIPE_REFERENCE = PIPE_SPECKLE | PIPE_MEDIAN
PIPE_HOTPATH = 0

STAGES = {  # name -> (id, dtype, per-pixel depth: 1 or "D")
    "census_left": (0, np.uint32, 1), "census_right": (1, np.uint32, 1), "aggr": (2, np.uint16, "D"),
    "disp_left_wta": (3, np.float32, 1), "disp_right": (4, np.float32, 1), "disp_lr": (5, np.float32, 1),
    "disp_speckle": (6, np.float32, 1), "disp_final": (7, np.float32, 1),
}
STAGE_PATH_PLANE_0 = 16
STAGE_GREY_LEFT, STAGE_GREY_RIGHT = 9, 10
GREY_BOARD, GREY_STB = 0, 1     # (76R+150G+29B)>>8 (board) / (77R+150G+29B)>>8 (stb_image, the demo's loader)


class SGMOption(C.Structure):
    """ABI of the reference's option struct (SemiGlobalMatching.h:24-40): sizeof 28, align 4."""
    _fields_ = [("num_paths", C.c_uint8), ("min_disparity", C.c_uint16), ("max_disparity", C.c_uint16),
                ("is_check_unique", C.c_bool), ("uniqueness_ratio", C.c_float), ("is_check_lr", C.c_bool),
                ("lrcheck_thres", C.c_float), ("is_remove_speckles", C.c_bool), ("min_speckle_area", C.c_uint16),
                ("p1", C.c_int16), ("p2_init", C.c_int16)]


def default_option(**kw) -> SGMOption:
    """The options the reference demo uses (main.c:48-65), overridable by keyword."""
    o = dict(num_paths=8, min_disparity=0, max_disparity=64, is_check_unique=True, uniqueness_ratio=0.99,
             is_check_lr=True, lrcheck_thres=1.0, is_remove_speckles=True, min_speckle_area=50, p1=10, p2_init=150)
    bad = set(kw) - set(o)
    if bad:
        raise KeyError(f"unknown SGMOption fields: {sorted(bad)}")
    o.update(kw)
    return SGMOption(**o)


def _load() -> C.CDLL:
    from . import build as _build      # imported lazily so `python -m <package>.build` runs cleanly
    path = _build.LIB
    if not os.path.isfile(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m soc_project_stereo_matching_b200.build` "
            "(needs nvcc; there is no CPU fallback for this library)")
    lib = C.CDLL(path)
    vp, u16, i32 = C.c_void_p, C.c_uint16, C.c_int
    sig = {
        "SGM_Initialize": (C.c_bool, [u16, u16, C.POINTER(SGMOption)]),
        "SGM_Reset": (C.c_bool, [u16, u16, C.POINTER(SGMOption)]),
        "SGM_Match": (C.c_bool, [vp, vp, vp]),
        "SGMB_LastError": (C.c_char_p, []),
        "SGMB_DeviceCount": (i32, []),
        "SGMB_Create": (i32, [C.POINTER(vp), i32, i32]),
        "SGMB_Destroy": (None, [vp]),
        "SGMB_Configure": (i32, [vp, u16, u16, C.POINTER(SGMOption)]),
        "SGMB_SetPipeline": (i32, [vp, C.c_uint]),
        "SGMB_SetCensusWindow": (i32, [vp, i32, i32]),
        "SGMB_SetGlobalCensusWindow": (i32, [i32, i32]),
        "SGMB_Match": (i32, [vp, vp, vp, vp]),
        "SGMB_MatchDevice": (i32, [vp, vp, vp, vp, i32]),
        "SGMB_Synchronize": (i32, [vp]),
        "SGMB_MatchBatch": (i32, [vp, vp, vp, vp, i32]),
        "SGMB_MatchBatchDevice": (i32, [vp, vp, vp, vp, i32]),
        "SGMB_MatchBatchMultiGPU": (i32, [vp, i32, i32, u16, u16, C.POINTER(SGMOption), C.c_uint, vp, vp, vp, i32]),
        "SGMB_GetStage": (i32, [vp, i32, vp, C.c_size_t]),
        "SGMB_PoolCreate": (i32, [C.POINTER(vp), vp, i32, i32]),
        "SGMB_PoolDestroy": (None, [vp]),
        "SGMB_PoolSize": (i32, [vp]),
        "SGMB_PoolContext": (vp, [vp, i32]),
        "SGMB_PoolConfigure": (i32, [vp, u16, u16, C.POINTER(SGMOption), C.c_uint]),
        "SGMB_PoolMatchBatch": (i32, [vp, vp, vp, vp, i32]),
        "SGMB_SetGreyFormula": (i32, [vp, i32]),
        "SGMB_MatchFrame": (i32, [vp, vp, vp, vp]),
        "SGMB_MatchFrameDevice": (i32, [vp, vp, vp, vp, i32]),
        "SGMB_DepthReplyBytes": (C.c_size_t, [u16, u16]),
        "SGMB_PackDepthReply": (i32, [C.c_uint32, u16, u16, vp, vp, C.c_size_t]),
        "SGMB_ParseFrameHeader": (i32, [vp, C.POINTER(i32), C.POINTER(C.c_int32), C.POINTER(u16), C.POINTER(u16), C.POINTER(C.c_size_t)]),
        "SGMB_DisparityToDepth": (i32, [vp, vp, C.c_size_t, C.c_float, C.c_float, C.c_float, vp]),
        "SGMB_DisparityToDepthDevice": (i32, [vp, vp, C.c_size_t, C.c_float, C.c_float, C.c_float, vp]),
        "SGMB_CompareDepth": (i32, [vp, vp, vp, C.c_size_t, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
        "SGMB_CompareDepthDevice": (i32, [vp, vp, vp, C.c_size_t, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
        "SGMB_HostAlloc": (i32, [C.POINTER(vp), C.c_size_t]),
        "SGMB_HostFree": (None, [vp]),
        "SGMB_HostRegister": (i32, [vp, C.c_size_t]),
        "SGMB_HostUnregister": (i32, [vp]),
        "SGMB_KernelLaunchesPerFrame": (i32, [vp]),
        "SGMB_ModelBytesPerFrame": (C.c_double, [vp]),
        "SGMB_PlanBytesPerFrame": (C.c_double, [vp]),
        "SGMB_LastDeviceMs": (C.c_float, [vp]),
        "SGMB_TimeDevice": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
        "SGMB_RunDevice": (i32, [vp, vp, vp, vp, i32, vp, vp]),
        "SGMB_RunDeviceReplays": (i32, [vp, vp, vp, vp, i32, i32, vp, vp]),
        "SGMB_TimeKernels": (i32, [vp, vp, vp, vp, i32, i32, vp, i32]),
        "SGMB_KernelName": (C.c_char_p, [vp, i32]),
        "SGMB_ShardRange": (i32, [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]),
        "SGMB_GlobalContext": (vp, []),
        "SGMB_SetGlobalDevice": (i32, [i32]),
        "SGMB_DebugWalkPath": (i32, [i32, i32, i32, i32, vp, i32]),
        "SGMB_DebugClassifyPaths": (i32, [i32, i32, i32, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


class _LazyLib:
    """Loads libsgm_b200.so on first use, so that `python -m soc_project_stereo_matching_b200.build` can
    import this package before the library exists.  Any use without the library raises ImportError."""
    _lib = None

    def __getattr__(self, name):
        if _LazyLib._lib is None:
            _LazyLib._lib = _load()
        return getattr(_LazyLib._lib, name)


lib = _LazyLib()


def last_error() -> str:
    return lib.SGMB_LastError().decode()


def _img_ptr(a, w=None, h=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- the reference's three functions
_state = {"shape": None}


def SGM_Initialize(width: int, height: int, option: SGMOption | None) -> bool:
    """bool SGM_Initialize(uint16_t, uint16_t, const SGMOption*) -- SemiGlobalMatching.h:78."""
    ok = bool(lib.SGM_Initialize(width, height, C.byref(option) if option is not None else None))
    _state["shape"] = (height, width) if ok else None
    return ok


def SGM_Reset(width: int, height: int, option: SGMOption | None) -> bool:
    """bool SGM_Reset(uint16_t, uint16_t, const SGMOption*) -- SemiGlobalMatching.h:79."""
    ok = bool(lib.SGM_Reset(width, height, C.byref(option) if option is not None else None))
    _state["shape"] = (height, width) if ok else None
    return ok


def SGM_Match(img_left: np.ndarray | None, img_right: np.ndarray | None, disp_left: np.ndarray) -> bool:
    """bool SGM_Match(const uint8_t*, const uint8_t*, float*) -- SemiGlobalMatching.h:80.

    ``disp_left`` must be a C-contiguous float32 array of width*height elements; it is overwritten."""
    if disp_left.dtype != np.float32 or not disp_left.flags.c_contiguous:
        raise TypeError("disp_left must be a C-contiguous float32 array")
    if _state["shape"] is not None and disp_left.size != _state["shape"][0] * _state["shape"][1]:
        raise ValueError("disp_left must hold width*height floats (the library writes that many)")
    keep = []
    ptrs = []
    for img in (img_left, img_right):
        if img is None:
            ptrs.append(None)
        else:
            a = np.ascontiguousarray(img, dtype=np.uint8)
            if _state["shape"] is not None and a.size != _state["shape"][0] * _state["shape"][1]:
                raise ValueError("image size does not match SGM_Initialize")
            keep.append(a)
            ptrs.append(a.ctypes.data_as(C.c_void_p))
    return bool(lib.SGM_Match(ptrs[0], ptrs[1], disp_left.ctypes.data_as(C.c_void_p)))


# ---------------------------------------------------------------- explicit contexts (sgm_b200.h)
class SGMError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"libsgm_b200 error {code}: {last_error()}")
        self.code = code


def _check(rc: int) -> None:
    if rc != 0:
        raise SGMError(rc)


class Context:
    """One SGMB_Context: a device, a configuration and ``slots`` frames in flight."""

    def __init__(self, device: int = 0, slots: int = 1):
        self._h = C.c_void_p()
        _check(lib.SGMB_Create(C.byref(self._h), device, slots))
        self.device, self.slots = device, slots
        self.width = self.height = self.disp_range = 0
        self.option = None
        self.census_window = (5, 5)

    def close(self) -> None:
        if self._h:
            lib.SGMB_Destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_pipeline(self, flags: int) -> None:
        _check(lib.SGMB_SetPipeline(self._h, flags))

    def set_census_window(self, width: int, height: int) -> None:
        """5x5 (reference, uint32 descriptors) or 9x7 (64-bit extension); applies at the next configure()."""
        _check(lib.SGMB_SetCensusWindow(self._h, width, height))
        self.census_window = (width, height)

    def configure(self, width: int, height: int, option: SGMOption) -> None:
        _check(lib.SGMB_Configure(self._h, width, height, C.byref(option)))
        self.width, self.height, self.option = width, height, option
        self.disp_range = option.max_disparity - option.min_disparity

    def match(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        l = np.ascontiguousarray(left, np.uint8); r = np.ascontiguousarray(right, np.uint8)
        if l.shape != (self.height, self.width) or r.shape != l.shape:
            raise ValueError("image shape does not match the configuration")
        out = np.empty((self.height, self.width), np.float32)
        _check(lib.SGMB_Match(self._h, l.ctypes.data, r.ctypes.data, out.ctypes.data))
        return out

    def match_ptr(self, left_ptr: int, right_ptr: int, out_ptr: int) -> None:
        """Host pointers (e.g. pinned buffers), blocking."""
        _check(lib.SGMB_Match(self._h, left_ptr, right_ptr, out_ptr))

    def match_device(self, d_left: int, d_right: int, d_out: int, sync: bool = True) -> None:
        _check(lib.SGMB_MatchDevice(self._h, d_left, d_right, d_out, int(sync)))

    def synchronize(self) -> None:
        _check(lib.SGMB_Synchronize(self._h))

    @staticmethod
    def _ptr_array(ptrs):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        return arr

    def match_batch_ptrs(self, lefts, rights, outs, device_memory: bool = False) -> None:
        n = len(lefts)
        fn = lib.SGMB_MatchBatchDevice if device_memory else lib.SGMB_MatchBatch
        _check(fn(self._h, self._ptr_array(lefts), self._ptr_array(rights), self._ptr_array(outs), n))

    def match_batch(self, lefts: np.ndarray, rights: np.ndarray) -> np.ndarray:
        """lefts/rights: uint8 [n, H, W] -> float32 [n, H, W]."""
        l, r = _batch_arrays(lefts, rights, self.height, self.width)
        n = l.shape[0]
        out = np.empty((n, self.height, self.width), np.float32)
        self.match_batch_ptrs([l[k].ctypes.data for k in range(n)], [r[k].ctypes.data for k in range(n)],
                              [out[k].ctypes.data for k in range(n)])
        return out

    def stage(self, name: str) -> np.ndarray:
        sid, dtype, depth = STAGES[name]
        if name.startswith("census") and self.census_window == (9, 7):
            dtype = np.uint64
        shape = (self.height, self.width) + ((self.disp_range,) if depth == "D" else ())
        out = np.empty(shape, dtype)
        _check(lib.SGMB_GetStage(self._h, sid, out.ctypes.data, out.nbytes))
        return out

    # ---- frame formats either side of the path and evaluation (N4)
    def set_grey_formula(self, formula: int) -> None:
        _check(lib.SGMB_SetGreyFormula(self._h, formula))

    def match_frame(self, planes6: np.ndarray, calib20: np.ndarray | None = None) -> np.ndarray:
        """planes6: uint8 [6, H, W] = left B,G,R then right B,G,R (the board's frame layout).  Returns the
        disparity map, or the depth map when the 20-float wire calibration is given."""
        p = np.ascontiguousarray(planes6, np.uint8)
        if p.shape != (6, self.height, self.width):
            raise ValueError("planes6 must have shape (6, H, W)")
        cal = None if calib20 is None else np.ascontiguousarray(calib20, np.float32)
        if cal is not None and cal.size != 20:
            raise ValueError("calib20 must hold 20 floats")
        out = np.empty((self.height, self.width), np.float32)
        _check(lib.SGMB_MatchFrame(self._h, p.ctypes.data, cal.ctypes.data if cal is not None else None, out.ctypes.data))
        return out

    def grey(self, right: bool = False) -> np.ndarray:
        out = np.empty((self.height, self.width), np.uint8)
        _check(lib.SGMB_GetStage(self._h, STAGE_GREY_RIGHT if right else STAGE_GREY_LEFT, out.ctypes.data, out.nbytes))
        return out

    def disparity_to_depth(self, disp: np.ndarray, baseline: float, fx: float, doffs: float) -> np.ndarray:
        d = np.ascontiguousarray(disp, np.float32)
        out = np.empty_like(d)
        _check(lib.SGMB_DisparityToDepth(self._h, d.ctypes.data, d.size, baseline, fx, doffs, out.ctypes.data))
        return out

    def compare_depth(self, ground_truth: np.ndarray, test: np.ndarray, abs_thresh: float = 10.0):
        """-> (rmse, bad-pixel rate, n_valid) like HostScript_Server/depth_image.py compare_img."""
        g = np.ascontiguousarray(ground_truth, np.float32); t = np.ascontiguousarray(test, np.float32)
        if g.shape != t.shape:
            raise ValueError("shape mismatch")
        rmse, bpr, nv = C.c_double(), C.c_double(), C.c_longlong()
        _check(lib.SGMB_CompareDepth(self._h, g.ctypes.data, t.ctypes.data, g.size, abs_thresh, C.byref(rmse), C.byref(bpr), C.byref(nv)))
        return rmse.value, bpr.value, nv.value

    def speckle_labels(self):
        """Debug tap: (root per pixel or -1, size per root) of the speckle filter's last run."""
        out = np.empty((2, self.height, self.width), np.int32)
        _check(lib.SGMB_GetStage(self._h, 8, out.ctypes.data, out.nbytes))
        return out[0], out[1]

    def path_plane(self, direction: int) -> np.ndarray:
        out = np.empty((self.height, self.width, self.disp_range), np.uint8)
        _check(lib.SGMB_GetStage(self._h, STAGE_PATH_PLANE_0 + direction, out.ctypes.data, out.nbytes))
        return out

    def kernel_launches_per_frame(self) -> int:
        return lib.SGMB_KernelLaunchesPerFrame(self._h)

    def model_bytes_per_frame(self) -> float:
        return lib.SGMB_ModelBytesPerFrame(self._h)

    def plan_bytes_per_frame(self) -> float:
        return lib.SGMB_PlanBytesPerFrame(self._h)

    def last_device_ms(self) -> float:
        return lib.SGMB_LastDeviceMs(self._h)

    def time_device(self, d_left: int, d_right: int, d_out: int, warmup: int, iters: int, flush_l2: bool = True):
        frame = np.zeros(iters, np.float32); agg = np.zeros(iters, np.float32)
        _check(lib.SGMB_TimeDevice(self._h, d_left, d_right, d_out, warmup, iters, int(flush_l2), frame.ctypes.data, agg.ctypes.data))
        return frame, agg


    def run_device_replays(self, d_left: int, d_right: int, d_out: int, iters: int, replays: int):
        """-> (milliseconds of each of `replays` launches of one graph of `iters` frames, aggregation-kernel ms of the last)."""
        rep = np.zeros(replays, np.float32); agg = np.zeros(iters, np.float32)
        _check(lib.SGMB_RunDeviceReplays(self._h, d_left, d_right, d_out, iters, replays, rep.ctypes.data, agg.ctypes.data))
        return rep, agg

    def time_kernels(self, d_left: int, d_right: int, d_out: int, warmup: int = 3, iters: int = 20) -> list[tuple[str, float]]:
        """-> [(kernel name, mean milliseconds)] of one frame of the current pipeline (CUDA events around every launch)."""
        ms = np.zeros(16, np.float32)
        n = lib.SGMB_TimeKernels(self._h, d_left, d_right, d_out, warmup, iters, ms.ctypes.data, ms.size)
        if n < 0:
            raise SGMError(n)
        return [(lib.SGMB_KernelName(self._h, k).decode(), float(ms[k])) for k in range(n)]

    def run_device(self, d_left: int, d_right: int, d_out: int, iters: int, time_aggregation: bool = True):
        """-> (total milliseconds of `iters` back-to-back frames, per-launch aggregation-kernel ms)."""
        total = C.c_float(0)
        agg = np.zeros(iters, np.float32)
        _check(lib.SGMB_RunDevice(self._h, d_left, d_right, d_out, iters, C.byref(total),
                                  agg.ctypes.data if time_aggregation else None))
        return float(total.value), agg


def shard_range(n: int, ndev: int, g: int) -> tuple[int, int]:
    """The library's batch sharding rule (SGMB_ShardRange, host only): device g of ndev takes pairs [lo, hi) of n."""
    lo, hi = C.c_int(), C.c_int()
    _check(lib.SGMB_ShardRange(n, ndev, g, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def _batch_arrays(lefts, rights, height: int, width: int):
    l = np.ascontiguousarray(lefts, np.uint8); r = np.ascontiguousarray(rights, np.uint8)
    if l.ndim != 3 or l.shape != r.shape or l.shape[1:] != (height, width):
        raise ValueError(f"lefts / rights must both have shape (n, {height}, {width}); got {l.shape} and {r.shape}")
    return l, r


def match_batch_multi_gpu(devices, slots_per_device, width, height, option, pipeline, lefts, rights) -> np.ndarray:
    l, r = _batch_arrays(lefts, rights, height, width)
    n = l.shape[0]
    out = np.empty((n, height, width), np.float32)
    dev = (C.c_int * len(devices))(*devices)
    pa = lambda a: (C.c_void_p * n)(*[a[k].ctypes.data for k in range(n)])
    _check(lib.SGMB_MatchBatchMultiGPU(dev, len(devices), slots_per_device, width, height, C.byref(option), pipeline,
                                       pa(l), pa(r), pa(out), n))
    return out


class Pool:
    """Persistent multi-GPU pool: one context per device, batches sharded contiguously (pair k -> device k*ndev//n)."""

    def __init__(self, devices, slots_per_device: int = 2):
        self._h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        _check(lib.SGMB_PoolCreate(C.byref(self._h), dev, len(devices), slots_per_device))
        self.width = self.height = 0

    def close(self) -> None:
        if self._h:
            lib.SGMB_PoolDestroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self) -> int:
        return lib.SGMB_PoolSize(self._h)

    def configure(self, width: int, height: int, option: SGMOption, pipeline: int = PIPE_REFERENCE) -> None:
        _check(lib.SGMB_PoolConfigure(self._h, width, height, C.byref(option), pipeline))
        self.width, self.height = width, height

    def match_batch(self, lefts: np.ndarray, rights: np.ndarray) -> np.ndarray:
        l, r = _batch_arrays(lefts, rights, self.height, self.width)
        n = l.shape[0]
        out = np.empty((n, self.height, self.width), np.float32)
        pa = lambda a: (C.c_void_p * n)(*[a[k].ctypes.data for k in range(n)])
        _check(lib.SGMB_PoolMatchBatch(self._h, pa(l), pa(r), pa(out), n))
        return out


def debug_walk_path(width: int, height: int, direction: int, path: int) -> np.ndarray:
    """Pixel indices visited by one aggregation path according to the product's walker (host code)."""
    buf = np.empty(max(width, height), np.int32)
    n = lib.SGMB_DebugWalkPath(width, height, direction, path, buf.ctypes.data, buf.size)
    if n < 0:
        raise SGMError(n)
    return buf[:n].copy()


def debug_classify_paths(width: int, height: int, direction: int) -> np.ndarray:
    """uint8[npaths]: 1 where the product treats the path of ``direction`` as irregular."""
    npaths = height if direction < 2 else width
    buf = np.zeros(npaths, np.uint8)
    n = lib.SGMB_DebugClassifyPaths(width, height, direction, buf.ctypes.data, buf.size)
    if n < 0:
        raise SGMError(n)
    return buf


def pack_depth_reply(frame_id: int, depth: np.ndarray) -> bytes:
    """The board's reply message (zb/tcp_perf_client.c:106-131): b'\x03' + '<IHH' + float32 rows."""
    d = np.ascontiguousarray(depth, np.float32)
    h, w = d.shape
    n = lib.SGMB_DepthReplyBytes(w, h)
    buf = (C.c_uint8 * n)()
    _check(lib.SGMB_PackDepthReply(frame_id, w, h, d.ctypes.data, buf, n))
    return bytes(buf)


def parse_frame_header(header9: bytes):
    """-> (type, seq, width, height, payload_bytes) of the server's '<BiHH' frame header (server.py:114)."""
    b = (C.c_uint8 * 9)(*header9[:9])
    t, q, w, h, pb = C.c_int(), C.c_int32(), C.c_uint16(), C.c_uint16(), C.c_size_t()
    _check(lib.SGMB_ParseFrameHeader(b, C.byref(t), C.byref(q), C.byref(w), C.byref(h), C.byref(pb)))
    return t.value, q.value, w.value, h.value, pb.value

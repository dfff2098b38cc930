"""GPU: the drop-in boundary exercised the way the reference's own callers use it.

* the reference's UNMODIFIED demo driver (main.c) linked against libsgm_b200.so instead of SemiGlobalMatching.c,
  EXECUTED on the GPU box (staged by oracle/build_ref.py::build_demo into oracle/_ref/demo/), its output PNG compared
  with the reference tree's own known-answer artefact Data/cone/im2.d.png (tests/golden/cone_demo.npz);
* host-memory kinds: pageable (malloc'd / static arrays, what main.c passes) and page-locked buffers give the same bits;
* two live contexts with different disparity ranges on one device (shared per-function attributes);
* error paths of the batch entry point leave nothing in flight."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import build_ref
import soc_project_stereo_matching_b200 as sgm
from helpers import GOLDEN, assert_same, load_golden, to_sgm_option
from pyoracle import options
from soc_project_stereo_matching_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def _read_png_grey(path: str) -> np.ndarray:
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert img is not None, path
        return img
    except ImportError:
        from PIL import Image
        return np.asarray(Image.open(path))


def test_reference_demo_runs_against_the_library_and_reproduces_its_png():
    """SURVEY.md section 7.2 acceptance run (main.c:16-120): load ../Data/cone/im2.png + im6.png with stb_image,
    SGM_Initialize + SGM_Match, min-max normalise, write ../Data/cone/im2.d.png."""
    exe = build_ref.build_demo()
    if not exe or not os.path.isfile(exe):
        pytest.skip("demo binary not staged (run __graft_entry__.build() where /root/reference exists)")
    out_png = os.path.join(os.path.dirname(exe), "..", "Data", "cone", "im2.d.png")
    if os.path.exists(out_png):
        os.remove(out_png)
    res = subprocess.run([exe], cwd=os.path.dirname(exe), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Saving Disparity Map" in res.stdout
    got = _read_png_grey(out_png)
    demo = np.load(os.path.join(GOLDEN, "cone_demo.npz"))["demo"]
    assert got.shape == demo.shape == (375, 450)
    # the reference binary itself reproduces its committed PNG on all but one pixel
    assert int((got != demo).sum()) <= 1, f"{int((got != demo).sum())} pixels differ from Data/cone/im2.d.png"
    # and bit-exactly what main.c:92-120 makes of the reference's own final disparity map
    _, _, _, want = load_golden("cone")
    d = want["disp_final"]
    v = np.isfinite(d)
    mn, mx = d[v].min(), d[v].max()
    img = np.zeros(d.shape, np.uint8)
    img[v] = np.clip((d[v] - mn) / (mx - mn) * np.float32(255.0), 0, 255).astype(np.uint8)
    assert np.array_equal(got, img)


def test_pinned_and_pageable_host_buffers_give_the_same_result(oracle):
    w, h, d = 200, 60, 64
    opts = options(max_disparity=d)
    left, right, _ = make_pair(w, h, d, seed=77, texture="scene")
    want = oracle.match(left, right, opts, stages=False)["disp_final"]
    n = w * h
    with sgm.Context(0, slots=2) as c:
        c.configure(w, h, to_sgm_option(opts))
        assert_same("pageable", c.match(left, right), want)                       # numpy arrays are pageable memory
        ptrs = []
        try:
            for nbytes in (n, n, 4 * n):
                p = C.c_void_p()
                assert sgm.lib.SGMB_HostAlloc(C.byref(p), nbytes) == 0, sgm.last_error()
                ptrs.append(p)
            C.memmove(ptrs[0], left.ctypes.data, n); C.memmove(ptrs[1], right.ctypes.data, n)
            c.match_ptr(ptrs[0].value, ptrs[1].value, ptrs[2].value)
            got = np.frombuffer((C.c_float * n).from_address(ptrs[2].value), np.float32).reshape(h, w).copy()
            assert_same("pinned", got, want)
            # mixed: pinned inputs, pageable output, and a batch that reuses its slots (pending pageable outputs are flushed)
            out = np.zeros((5, h, w), np.float32)
            c.match_batch_ptrs([ptrs[0].value] * 5, [ptrs[1].value] * 5, [out[k].ctypes.data for k in range(5)])
            for k in range(5):
                assert_same(f"batch[{k}] pageable out", out[k], want)
        finally:
            for p in ptrs:
                sgm.lib.SGMB_HostFree(p)


def test_registered_caller_memory_and_direct_output(oracle):
    """Caller-owned arrays page-locked with SGMB_HostRegister: the last kernel writes the output buffer directly (no copy at
    the end of the call).  The buffer is pre-filled with garbage, several calls alternate between two output buffers and a
    pageable one (the recorded graph's output argument is re-pointed each time), and a batch writes five registered outputs."""
    w, h, d = 333, 75, 96
    opts = options(max_disparity=d)
    pairs = [make_pair(w, h, d, seed=500 + k, texture="scene")[:2] for k in range(2)]
    wants = [oracle.match(l, r, opts, stages=False)["disp_final"] for l, r in pairs]
    outs = [np.full((h, w), -7.0, np.float32) for _ in range(2)]
    batch_out = np.full((5, h, w), -3.0, np.float32)
    pageable = np.full((h, w), -5.0, np.float32)
    reg = [a for pr in pairs for a in pr] + outs + [batch_out]
    for a in reg:
        assert sgm.lib.SGMB_HostRegister(a.ctypes.data, a.nbytes) == 0, sgm.last_error()
    try:
        with sgm.Context(0, slots=2) as c:
            c.configure(w, h, to_sgm_option(opts))
            for rep in range(6):
                k = rep % 2
                (l, r), dst = pairs[k], (pageable if rep == 3 else outs[(rep // 2) % 2])
                dst[:] = -9.0
                c.match_ptr(l.ctypes.data, r.ctypes.data, dst.ctypes.data)
                assert_same(f"registered rep {rep}", dst, wants[k])
            c.match_batch_ptrs([pairs[k % 2][0].ctypes.data for k in range(5)], [pairs[k % 2][1].ctypes.data for k in range(5)],
                               [batch_out[k].ctypes.data for k in range(5)])
            for k in range(5):
                assert_same(f"registered batch[{k}]", batch_out[k], wants[k % 2])
            with sgm.Context(0, slots=1) as one:             # a single-slot context writes batch results directly as well
                one.configure(w, h, to_sgm_option(opts))
                batch_out[:] = -3.0
                one.match_batch_ptrs([pairs[k % 2][0].ctypes.data for k in range(5)], [pairs[k % 2][1].ctypes.data for k in range(5)],
                                     [batch_out[k].ctypes.data for k in range(5)])
                for k in range(5):
                    assert_same(f"registered batch, one slot [{k}]", batch_out[k], wants[k % 2])
            # with the taps on the result also has to stay readable as a stage: the copy path is used
            c.set_pipeline(sgm.PIPE_REFERENCE | sgm.PIPE_TAPS)
            c.configure(w, h, to_sgm_option(opts))
            outs[0][:] = -9.0
            c.match_ptr(pairs[0][0].ctypes.data, pairs[0][1].ctypes.data, outs[0].ctypes.data)
            assert_same("registered, taps on", outs[0], wants[0])
            assert_same("stage disp_final", c.stage("disp_final"), wants[0])
    finally:
        for a in reg:
            assert sgm.lib.SGMB_HostUnregister(a.ctypes.data) == 0, sgm.last_error()
    assert sgm.lib.SGMB_HostRegister(None, 16) != 0 and sgm.lib.SGMB_HostUnregister(None) != 0


def test_two_live_contexts_with_different_disparity_ranges(oracle):
    """The opt-in shared-memory size of the WTA kernel is a per-function attribute shared by all contexts of the process:
    configuring a second context with a smaller range must not break the first one (D=128 then D=112; D=256 then D=200)."""
    for d_big, d_small in ((128, 112), (256, 200)):
        w, h = 300, 20
        la, ra, _ = make_pair(w, h, d_big, seed=1, texture="scene")
        lb, rb, _ = make_pair(w, h, d_small, seed=2, texture="scene")
        oa, ob = options(max_disparity=d_big), options(max_disparity=d_small)
        with sgm.Context(0) as a, sgm.Context(0) as b:
            a.configure(w, h, to_sgm_option(oa))
            first = a.match(la, ra)
            b.configure(w, h, to_sgm_option(ob))
            assert_same("small", b.match(lb, rb), oracle.match(lb, rb, ob, stages=False)["disp_final"])
            a.set_pipeline(sgm.PIPE_HOTPATH)            # drops the recorded frame graph: the kernels are launched afresh
            a.set_pipeline(sgm.PIPE_REFERENCE)
            assert_same("big after small", a.match(la, ra), first)
            assert_same("big", first, oracle.match(la, ra, oa, stages=False)["disp_final"])


def test_batch_rejects_null_pointers_before_enqueuing_anything():
    w, h, d = 64, 32, 16
    left, right, _ = make_pair(w, h, d, seed=3, texture="noise")
    out = np.full((3, h, w), -1.0, np.float32)
    with sgm.Context(0, slots=2) as c:
        c.configure(w, h, sgm.default_option(max_disparity=d))
        with pytest.raises(sgm.SGMError):
            c.match_batch_ptrs([left.ctypes.data] * 3, [right.ctypes.data, None, right.ctypes.data], [out[k].ctypes.data for k in range(3)])
        assert np.all(out == -1.0), "a pair was processed although the batch was rejected"
        got = c.match_batch(np.stack([left] * 3), np.stack([right] * 3))      # the context is still usable
        assert np.array_equal(got[0].view(np.uint32), got[2].view(np.uint32))


def test_multi_gpu_pool_shards_match_the_oracle(oracle):
    """SGMB_Pool* over every visible device with more pairs than devices; needs at least two GPUs (the single-GPU box
    runs the same entry points with ndev == 1 in test_parity_gpu.py)."""
    ndev = sgm.lib.SGMB_DeviceCount()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    w, h, d, n = 160, 48, 64, 4 * ndev + 3
    opts = options(max_disparity=d)
    pairs = [make_pair(w, h, d, seed=900 + k, texture="scene" if k % 2 else "noise")[:2] for k in range(n)]
    lefts = np.stack([p[0] for p in pairs]); rights = np.stack([p[1] for p in pairs])
    with sgm.Pool(list(range(ndev)), slots_per_device=3) as pool:
        pool.configure(w, h, to_sgm_option(opts), sgm.PIPE_REFERENCE)
        got = pool.match_batch(lefts, rights)
    for k in range(n):
        assert_same(f"pool over {ndev} GPUs [{k}]", got[k], oracle.match(lefts[k], rights[k], opts, stages=False)["disp_final"])
    got2 = sgm.match_batch_multi_gpu(list(range(ndev)), 2, w, h, to_sgm_option(opts), sgm.PIPE_REFERENCE, lefts, rights)
    assert np.array_equal(got.view(np.uint32), got2.view(np.uint32))

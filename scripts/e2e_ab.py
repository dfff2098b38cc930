#!/usr/bin/env python3
"""A/B helper: SGM_Match at C2 with page-locked buffers, 300 calls, median / p95 in ms (run once per environment setting).
Usage: [SGM_B200_...=1] python scripts/e2e_ab.py [tag]"""
import ctypes as C, hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair
w, h, d = 1242, 375, 128
left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
assert sgm.SGM_Initialize(w, h, sgm.default_option(max_disparity=d)), sgm.last_error()
ptrs = []
for nb in (w * h, w * h, 4 * w * h):
    p = C.c_void_p(); assert sgm.lib.SGMB_HostAlloc(C.byref(p), nb) == 0; ptrs.append(p)
C.memmove(ptrs[0], left.ctypes.data, w * h); C.memmove(ptrs[1], right.ctypes.data, w * h)
call = lambda: sgm.lib.SGM_Match(ptrs[0], ptrs[1], ptrs[2])
for _ in range(20): assert call()
ts = []
for _ in range(300):
    t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
res = np.ctypeslib.as_array(C.cast(ptrs[2], C.POINTER(C.c_float)), shape=(h, w)).copy()
print(sys.argv[1] if len(sys.argv) > 1 else "", "median_ms %.4f p95_ms %.4f" % (np.median(ts) * 1e3, np.percentile(ts, 95) * 1e3), hashlib.md5(res.tobytes()).hexdigest())

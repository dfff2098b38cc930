#!/usr/bin/env python3
"""Per-kernel CUDA-event timings of one frame for several configurations, and SGM_Match end to end for the host-memory
kinds a caller can pass (page-locked, pageable staged by the library, pageable handed to the driver).
Usage: python scripts/prof_kernels.py [configs...]   (default: c2 c2p4 c1 c3)  -> JSON lines"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair

CONFIGS = {"c1": (450, 375, 64, 8), "c2": (1242, 375, 128, 8), "c2p4": (1242, 375, 128, 4), "c3": (2864, 1924, 256, 8),
           "c5": (3840, 2160, 256, 8)}


def dev_alloc(nbytes):
    import torch
    return torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def kernels(name):
    import torch
    w, h, d, paths = CONFIGS[name]
    left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
    dl = torch.from_numpy(left).cuda(); dr = torch.from_numpy(right).cuda()
    do = torch.empty((h, w), dtype=torch.float32, device="cuda")
    opt = sgm.default_option(max_disparity=d, num_paths=paths)
    with sgm.Context(0) as ctx:
        ctx.set_pipeline(sgm.PIPE_REFERENCE)
        ctx.configure(w, h, opt)
        ks = ctx.time_kernels(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 3, 20 if w < 2000 else 5)
        ctx.set_pipeline(sgm.PIPE_HOTPATH)
        rep, agg = ctx.run_device_replays(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 10 if w < 2000 else 2, 5)
    print(json.dumps({"config": name, "shape": [w, h, d, paths], "kernels_us": {k: round(v * 1e3, 1) for k, v in ks},
                      "sum_us": round(sum(v for _, v in ks) * 1e3, 1),
                      "hot_ms_per_frame_graph": round(float(np.median(rep)) / (10 if w < 2000 else 2), 4),
                      "agg_ms": round(float(np.mean(agg)), 4)}), flush=True)


def e2e_kinds():
    w, h, d = 1242, 375, 128
    left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
    opt = sgm.default_option(max_disparity=d)
    n = w * h
    res = {}
    code = r'''
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, %r)
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair
w, h, d = 1242, 375, 128
left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
opt = sgm.default_option(max_disparity=d)
kind = sys.argv[1]
assert sgm.SGM_Initialize(w, h, opt)
if kind == "pinned":
    ptrs = []
    for nb in (w*h, w*h, 4*w*h):
        p = C.c_void_p(); assert sgm.lib.SGMB_HostAlloc(C.byref(p), nb) == 0; ptrs.append(p)
    C.memmove(ptrs[0], left.ctypes.data, w*h); C.memmove(ptrs[1], right.ctypes.data, w*h)
    call = lambda: sgm.lib.SGM_Match(ptrs[0], ptrs[1], ptrs[2])
else:
    out = np.zeros((h, w), np.float32)
    call = lambda: sgm.lib.SGM_Match(left.ctypes.data, right.ctypes.data, out.ctypes.data)
for _ in range(10): assert call()
ts = []
for _ in range(200):
    t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
print(np.median(ts) * 1e3, np.percentile(ts, 95) * 1e3)
''' % ROOT
    for kind, env in (("pinned", {}), ("pageable", {})):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", code, kind], capture_output=True, text=True, env=e)
        try:
            med, p95 = (float(x) for x in out.stdout.split())
            res[kind] = {"median_ms": round(med, 4), "p95_ms": round(p95, 4)}
        except ValueError:
            res[kind] = {"error": out.stderr[-400:]}
    print(json.dumps({"e2e_SGM_Match_C2": res}), flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:]:                       # ad-hoc shapes: WxHxDxP
        if a.count("x") == 3:
            CONFIGS[a] = tuple(int(v) for v in a.split("x"))
    names = [a for a in sys.argv[1:] if a in CONFIGS] or ["c2", "c2p4", "c1", "c3"]
    for nme in names:
        kernels(nme)
    if "--no-e2e" not in sys.argv:
        e2e_kinds()

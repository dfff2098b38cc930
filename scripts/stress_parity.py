#!/usr/bin/env python3
"""Repeat the full pipeline many times on a few inputs and report the first stage that ever differs
(hunting nondeterministic failures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import soc_project_stereo_matching_b200 as sgm
from helpers import load_golden, to_sgm_option
from pyoracle import Oracle, options
from soc_project_stereo_matching_b200.synth import make_pair

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
taps = not (len(sys.argv) > 2 and sys.argv[2] == "notaps")
orc = Oracle()
cases = []
l, r, o, want = load_golden("cone")
cases.append(("cone", l, r, o, orc.match(l, r, o)))
for (w, h, d, tex) in [(1242, 375, 128, "noise"), (1242, 375, 128, "scene"), (640, 200, 64, "scene")]:
    o = options(max_disparity=d)
    l, r, _ = make_pair(w, h, d, seed=0xB200, texture=tex)
    cases.append((f"{w}x{h}x{d}/{tex}", l, r, o, orc.match(l, r, o)))
stages = ["aggr", "disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final"] if taps else ["disp_lr", "disp_final"]
bad = {}
with sgm.Context(0) as ctx:
    ctx.set_pipeline(sgm.PIPE_REFERENCE | (sgm.PIPE_TAPS if taps else 0))
    for name, l, r, o, want in cases:
        ctx.configure(l.shape[1], l.shape[0], to_sgm_option(o))
        for it in range(reps):
            final = ctx.match(l, r)
            got = {k: ctx.stage(k) for k in stages[:-1]}
            got["disp_final"] = final
            for k in stages:
                if not np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)):
                    n = int((got[k].view(np.uint32 if got[k].dtype == np.float32 else got[k].dtype) != want[k].view(np.uint32 if want[k].dtype == np.float32 else want[k].dtype)).sum())
                    bad.setdefault((name, k), []).append((it, n))
                    break
        print(name, "done", flush=True)
print("FAILURES:", {k: v[:5] + [len(v)] for k, v in bad.items()} if bad else "none")

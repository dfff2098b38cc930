#!/usr/bin/env python3
"""Tiny driver for ncu: a few KITTI-shaped (C2) frames through the device-resident path, hot path first and
then the full SGM_Match pipeline.  Usage: python profiles/prof_frame.py [frames] [WxHxD] [paths]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w, h, d = (int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1242x375x128").split("x"))
paths = int(sys.argv[3]) if len(sys.argv) > 3 else 8
left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
opt = sgm.default_option(max_disparity=d, num_paths=paths)
with sgm.Context(0) as ctx:
    for flags in (sgm.PIPE_HOTPATH, sgm.PIPE_REFERENCE):
        ctx.set_pipeline(flags)
        ctx.configure(w, h, opt)
        for _ in range(frames):
            out = ctx.match(left, right)
        print("pipeline", flags, "last frame", round(ctx.last_device_ms(), 4), "ms; valid", float(np.isfinite(out).mean()))

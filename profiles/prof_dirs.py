#!/usr/bin/env python3
"""Profiling aid: time the aggregation kernel with only some directions enabled (results are not meaningful)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    import soc_project_stereo_matching_b200 as sgm
    from soc_project_stereo_matching_b200.synth import make_pair
    w, h, d = 1242, 375, 128
    left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
    dl = torch.from_numpy(left).cuda(); dr = torch.from_numpy(right).cuda(); do = torch.empty((h, w), dtype=torch.float32, device="cuda")
    with sgm.Context(0) as ctx:
        ctx.set_pipeline(sgm.PIPE_HOTPATH)
        ctx.configure(w, h, sgm.default_option(max_disparity=d))
        ctx.run_device(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 3, False)
        tot, agg = ctx.run_device(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 20, True)
        print(sys.argv[1], "frame_ms", round(tot / 20, 4), "agg_ms", round(float(agg.mean()), 4))
else:
    for mask in (sys.argv[2:] if False else os.environ.get("MASKS", "0x100 0x01 0x03 0x04 0x0c 0x30 0xfc 0xff").split()):
        env = dict(os.environ, SGM_B200_DEBUG_DIRMASK=mask)
        print(subprocess.run([sys.executable, __file__, mask], env=env, capture_output=True, text=True).stdout.strip())

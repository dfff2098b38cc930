#!/usr/bin/env python3
"""Device-resident hot-path time of every BASELINE.json config shape (C1..C5), one frame at a time, CUDA events on the
launching stream (SGMB_RunDevice).  Prints one JSON line per config.  Usage: python profiles/prof_configs.py [iters]"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
peak = 6547.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
CONFIGS = [("C1 cone-shaped 450x375 D=64 8 paths", 450, 375, 64, 8), ("C2 KITTI-shaped 1242x375 D=128 8 paths", 1242, 375, 128, 8),
           ("C3 Middlebury-full-shaped 2864x1924 D=256 8 paths", 2864, 1924, 256, 8), ("C4 KITTI-shaped 1242x375 D=128 4 paths (one frame)", 1242, 375, 128, 4),
           ("C5 4K 3840x2160 D=256 8 paths", 3840, 2160, 256, 8)]
for name, w, h, d, paths in CONFIGS:
    left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
    dl = torch.from_numpy(left).cuda(); dr = torch.from_numpy(right).cuda(); do = torch.empty((h, w), dtype=torch.float32, device="cuda")
    with sgm.Context(0) as ctx:
        ctx.set_pipeline(sgm.PIPE_HOTPATH)
        ctx.configure(w, h, sgm.default_option(max_disparity=d, num_paths=paths))
        ctx.run_device(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 3, False)
        tot, agg = ctx.run_device(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), iters, True)
        ms = tot / iters
        model = ctx.model_bytes_per_frame()
        print(json.dumps({"config": name, "ms_per_frame": round(ms, 4), "aggregation_ms": round(float(agg.mean()), 4), "MDE_per_s": round(w * h * d / ms / 1e3, 1),
                          "algorithmic_bytes": model, "frame_roofline_frac": round(model / (ms * 1e-3) / 1e9 / peak, 3),
                          "plan_bytes": ctx.plan_bytes_per_frame(), "valid": round(float(torch.isfinite(do).float().mean()), 4)}), flush=True)
    del dl, dr, do
    torch.cuda.empty_cache()

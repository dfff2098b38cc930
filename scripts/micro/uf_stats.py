#!/usr/bin/env python3
"""Union-find statistics of the speckle filter (library built with -DSGM_SPECKLE_DEBUG): finds, links walked, longest walk,
unions requested, failed compare-and-swaps - per frame, for the 8- and 4-path maps of the C2 pair."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair
w, h, d = 1242, 375, 128
left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
sgm.lib.SGMB_DebugUfStats.argtypes = [C.c_void_p, C.c_int]
for paths in (8, 4):
    with sgm.Context(0) as ctx:
        ctx.set_pipeline(sgm.PIPE_REFERENCE)
        ctx.configure(w, h, sgm.default_option(max_disparity=d, num_paths=paths))
        buf = np.zeros(8, np.uint64)
        sgm.lib.SGMB_DebugUfStats(buf.ctypes.data, 1)
        out = ctx.match(left, right)
        sgm.lib.SGMB_DebugUfStats(buf.ctypes.data, 1)
        f, hops, mx, un, fail = (int(v) for v in buf[:5])
        print(f"paths {paths}: valid {np.isfinite(out).mean():.3f} finds {f} links walked {hops} ({hops / max(f, 1):.2f} per find) longest {mx} unions {un} failed CAS {fail}")

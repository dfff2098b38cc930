import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import soc_project_stereo_matching_b200 as sgm
from helpers import load_golden, to_sgm_option
from scipy import ndimage
l, r, o, want0 = load_golden("cone")
h, w = l.shape
from pyoracle import Oracle
want = Oracle().match(l, r, o)
with sgm.Context(0) as ctx:
    ctx.set_pipeline(sgm.PIPE_REFERENCE | sgm.PIPE_TAPS)
    ctx.configure(w, h, to_sgm_option(o))
    good = None
    for it in range(300):
        ctx.match(l, r)
        sp = ctx.stage("disp_speckle")
        lab, size = ctx.speckle_labels()
        ok = np.array_equal(sp.view(np.uint32), want["disp_speckle"].view(np.uint32))
        if ok and good is None: good = (lab.copy(), size.copy())
        if not ok:
            bad = np.argwhere(sp.view(np.uint32) != want["disp_speckle"].view(np.uint32))
            y, x = bad[0]
            print("iter", it, "nbad", len(bad), "first", (y, x), "lab", lab[y, x], "size[root]", size.ravel()[lab[y, x]])
            if good is not None:
                gl, gs = good
                print("  good lab", gl[y, x], "good size", gs.ravel()[gl[y, x]])
                # is the partition the same?
                roots_bad = lab[tuple(bad.T)]
                print("  bad roots", np.unique(roots_bad), "good roots", np.unique(gl[tuple(bad.T)]))
                gr = np.unique(gl[tuple(bad.T)])[0]
                members = (gl == gr)
                print("  good comp size", members.sum(), "roots in bad run over those members", np.unique(lab[members], return_counts=True))
                print("  sizes in bad run for these roots", [int(size.ravel()[q]) for q in np.unique(lab[members])])
                print("  is every lab a root? ", np.all(lab.ravel()[lab[lab>=0]] == lab[lab>=0]))
            break
    else:
        print("no failure in 300")

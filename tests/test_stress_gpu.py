"""GPU: repeated runs hunting NONDETERMINISTIC failures (lost unions in the speckle filter's lock-free union-find, the
median wavefront's inter-CTA hand-over, atomics of the irregular paths).  compute-sanitizer is not usable on this pool,
so determinism under repetition and under concurrency (several frames in flight, so that kernels of other frames
compete for SMs with the median's spin-waiting CTAs) is the guard.  Bounded by a time budget; at least 200 frames."""
import time

import numpy as np

import pytest
import soc_project_stereo_matching_b200 as sgm
from helpers import load_golden, to_sgm_option
from pyoracle import options
from soc_project_stereo_matching_b200.synth import make_pair

pytestmark = pytest.mark.gpu

BUDGET_S = 150.0


def _same(a, b):
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_repeated_runs_are_bit_identical_to_the_oracle(oracle):
    t_end = time.time() + BUDGET_S
    cases = []
    l, r, o, _ = load_golden("cone")
    cases.append(("cone", l, r, o))
    for tex in ("noise", "scene"):
        l, r, _ = make_pair(1242, 375, 128, seed=0xB200, texture=tex)
        cases.append((f"C2/{tex}", l, r, options(max_disparity=128)))
    l, r, _ = make_pair(640, 200, 64, seed=0xB200, texture="scene")
    cases.append(("640x200x64", l, r, options(max_disparity=64)))
    wants = [oracle.match(l, r, o) for _, l, r, o in cases]
    frames, failures = 0, []

    # ---- single frames, every stage tapped (the first runs of every case) then final maps only
    with sgm.Context(0) as ctx:
        ctx.set_pipeline(sgm.PIPE_REFERENCE | sgm.PIPE_TAPS)
        for (name, l, r, o), want in zip(cases, wants):
            ctx.configure(l.shape[1], l.shape[0], to_sgm_option(o))
            for it in range(40):
                final = ctx.match(l, r)
                frames += 1
                stages = ["aggr", "disp_left_wta", "disp_right", "disp_lr", "disp_speckle"] if it < 6 else ["disp_lr"]
                for k in stages:
                    if not _same(ctx.stage(k), want[k]):
                        failures.append((name, it, k)); break
                else:
                    if not _same(final, want["disp_final"]):
                        failures.append((name, it, "disp_final"))
                if time.time() > t_end:
                    break

    # ---- four frames in flight: kernels of different frames overlap each other and the median's waiting CTAs
    t_end += 60.0
    for (name, l, r, o), want in zip(cases[:3], wants[:3]):
        with sgm.Context(0, slots=4) as ctx:
            ctx.set_pipeline(sgm.PIPE_REFERENCE)
            ctx.configure(l.shape[1], l.shape[0], to_sgm_option(o))
            lefts = np.stack([l] * 8); rights = np.stack([r] * 8)
            for it in range(8):
                got = ctx.match_batch(lefts, rights)
                frames += 8
                for k in range(8):
                    if not _same(got[k], want["disp_final"]):
                        failures.append((name + " batch", it, k))
                if time.time() > t_end:
                    break
    assert not failures, f"{len(failures)} nondeterministic mismatches in {frames} frames: {failures[:10]}"
    assert frames >= 200, f"time budget allowed only {frames} frames"

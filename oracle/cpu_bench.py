#!/usr/bin/env python3
"""Time the reference's CPU implementation of the path on this host's cores.  TEST/BENCH
INFRASTRUCTURE ONLY: run as a subprocess by bench.py's cpu_baseline leg and by `bench.py --impl reference`.

What runs: the reference's own SemiGlobalMatching.c, compiled verbatim per shape by oracle/build_ref.py
(kind "reference"; prebuilt into oracle/_ref/ by __graft_entry__.build() so it is available on the GPU box
where /root/reference does not exist); if no such library exists for the shape, our C restatement
oracle/sgm_oracle.c (kind "port").  The reference keeps all state in globals, so parallelism is one
PROCESS per core, each processing whole frames (paths span the image; a frame cannot be split).

One "step" = every worker process matching one frame concurrently (`--procs` frames per step).

Output: one JSON line {"kind", "cores", "frames_per_step", "steps", "warmup", "span", "step_seconds": [...],
"seconds_per_step", "frames_per_s", "mde_per_s", "single_frame_seconds"}.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, w, h, d, paths, span, kind, nsteps, barrier, out_q):
    import numpy as np
    import pyoracle
    from soc_project_stereo_matching_b200.synth import make_pair

    opts = pyoracle.options(max_disparity=d, num_paths=paths, remove_speckles=True)
    left, right, _ = make_pair(w, h, d, seed=0xB200 + rank, texture="noise")   # the same inputs as bench.py's GPU arm
    if kind == "reference":
        eng = pyoracle.Reference(w, h, d, "p4" if paths == 4 else "")
        run = (lambda: eng.hotpath(left, right, opts)) if span == "hot" else (lambda: eng.match_plain(left, right, opts))
    else:
        eng = pyoracle.Oracle()
        run = (lambda: eng.hotpath(left, right, opts)) if span == "hot" else (lambda: eng.match(left, right, opts, stages=False))
    times = []
    for _ in range(nsteps):
        barrier.wait()
        t0 = time.perf_counter()
        res = run()
        times.append(time.perf_counter() - t0)
        barrier.wait()
    out_q.put((rank, times, float(np.isfinite(res).mean())))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="1242x375x128")
    ap.add_argument("--paths", type=int, default=8)
    ap.add_argument("--span", choices=["hot", "full"], default="full")
    ap.add_argument("--procs", type=int, default=0, help="worker processes (0 = all cores)")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    w, h, d = (int(x) for x in a.shape.lower().split("x"))
    procs = a.procs or (os.cpu_count() or 1)

    import build_ref
    kind = "port"
    if h <= w:
        path = build_ref.build_ref(w, h, d, "p4" if a.paths == 4 else "")
        if path and os.path.isfile(path):
            kind = "reference"
    if kind == "port":
        build_ref.build_oracle()

    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs + 1)
    q = ctx.Queue()
    nsteps = a.steps + a.warmup
    ps = [ctx.Process(target=_worker, args=(r, w, h, d, a.paths, a.span, kind, nsteps, barrier, q)) for r in range(procs)]
    for p in ps:
        p.start()
    step_seconds = []
    for _ in range(nsteps):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        step_seconds.append(time.perf_counter() - t0)
    results = [q.get() for _ in ps]
    for p in ps:
        p.join()
    timed = step_seconds[a.warmup:]
    sec = sum(timed) / len(timed)
    single = min(min(t[a.warmup:]) for _, t, _ in results)
    print(json.dumps({
        "kind": kind, "cores": procs, "frames_per_step": procs, "steps": a.steps, "warmup": a.warmup, "span": a.span,
        "step_seconds": [round(x, 4) for x in timed], "seconds_per_step": sec,
        "frames_per_s": procs / sec, "mde_per_s": procs * w * h * d / sec / 1e6,
        "single_frame_seconds": single, "valid_fraction": results[0][2],
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Per-warp timeline of the aggregation kernel (library built with -DSGM_AGG_TRACE, see scripts/micro/build_variant.sh).
Prints, per SM, when its warps finish and what they were; and per direction the distribution of job durations.
Usage (GPU box): cp scripts/micro/libs/trace.so soc_project_stereo_matching_b200/lib/libsgm_b200.so; python scripts/micro/agg_trace.py"""
import collections, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import soc_project_stereo_matching_b200 as sgm
from soc_project_stereo_matching_b200.synth import make_pair

w, h, d = 1242, 375, 128
left, right, _ = make_pair(w, h, d, seed=0xB200, texture="noise")
dl = torch.from_numpy(left).cuda(); dr = torch.from_numpy(right).cuda(); do = torch.empty((h, w), dtype=torch.float32, device="cuda")
with sgm.Context(0) as ctx:
    ctx.set_pipeline(sgm.PIPE_HOTPATH)
    ctx.configure(w, h, sgm.default_option(max_disparity=d))
    os.environ["SGM_B200_NO_GRAPH"] = "1"
    for _ in range(3):
        tot, agg = ctx.run_device(dl.data_ptr(), dr.data_ptr(), do.data_ptr(), 1, True)
    n = 2245 + 8
    buf = np.zeros(n * 3, np.uint64)
    sgm.lib.SGMB_DebugAggTrace.argtypes = [C.c_void_p, C.c_int]
    assert sgm.lib.SGMB_DebugAggTrace(buf.ctypes.data, n) == 0
t = buf.reshape(n, 3)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
start = (t[:, 0] - t0).astype(np.float64) * 1e-3
end = (t[:, 1] - t0).astype(np.float64) * 1e-3
dirs = (t[:, 2] >> np.uint64(32)).astype(int); sm = (t[:, 2] & np.uint64(0xffffffff)).astype(int)
print("agg_ms", float(agg.mean()), "warps", len(t), "kernel span us", end.max())
for dd in sorted(set(dirs)):
    m = dirs == dd
    print(f"dir {dd:2d}: n={m.sum():4d} start {start[m].min():6.1f}..{start[m].max():6.1f}  dur min/med/max {np.min(end[m]-start[m]):6.1f} {np.median(end[m]-start[m]):6.1f} {np.max(end[m]-start[m]):6.1f}  end max {end[m].max():6.1f}")
per = collections.defaultdict(list)
for i in range(len(t)):
    per[sm[i]].append((end[i], dirs[i]))
rows = []
for s, v in per.items():
    cnt = collections.Counter(x[1] for x in v)
    nh = cnt[0] + cnt[1]; ni = sum(c for k, c in cnt.items() if k >= 8)
    rows.append((max(x[0] for x in v), s, len(v), nh, ni))
rows.sort()
print("SM finish times (us): min", rows[0][0], "median", rows[len(rows) // 2][0], "max", rows[-1][0])
hist = collections.Counter((r[2], r[3], r[4]) for r in rows)
for k in sorted(hist):
    sel = [r[0] for r in rows if (r[2], r[3], r[4]) == k]
    print(f"SMs with {k[0]:2d} warps ({k[1]} horizontal, {k[2]} irregular): {hist[k]:3d}  finish min/med/max {min(sel):6.1f} {np.median(sel):6.1f} {max(sel):6.1f}")
if os.environ.get("DUMP_PLACEMENT"):
    full = buf.reshape(n, 3)
    cta_sm = [(int(full[4 * i, 2] & np.uint64(0xffffffff)) if full[4 * i, 0] > 0 else -1) for i in range(n // 4)]
    print("cta->sm first 40:", cta_sm[:40])
    print("cta->sm 148..188:", cta_sm[148:188])
    same = sum(1 for i in range(len(cta_sm) - 148) if cta_sm[i] == cta_sm[i + 148])
    print("ctas with sm[i] == sm[i+148]:", same, "of", len(cta_sm) - 148)
    order = {}
    for i, s_ in enumerate(cta_sm):
        order.setdefault(s_, []).append(i)
    print("per-SM CTA lists (first 12 SMs by id):", [(k, order[k]) for k in sorted(order)[:12]])
if os.environ.get("DUMP_SLOWEST"):
    byfin = sorted(rows)
    def describe(r):
        s_ = r[1]
        cnt = collections.Counter(x[1] for x in per[s_])
        last = max(per[s_])
        return f"SM {s_:3d} finish {r[0]:6.1f} dirs {dict(sorted(cnt.items()))} last job dir {last[1]}"
    print("slowest SMs:"); [print("  ", describe(r)) for r in byfin[-10:]]
    hs = [r for r in byfin if r[3] == 4 and r[2] == 16]
    print("fastest SMs with 4 horizontal + 12 column-like:"); [print("  ", describe(r)) for r in hs[:6]]
    # when do the warps of each direction end on the slowest SM
    s_ = byfin[-1][1]
    print("jobs of the slowest SM:", sorted((round(e, 1), d_) for e, d_ in per[s_]))
    s_ = hs[0][1]
    print("jobs of the fastest 4+12 SM:", sorted((round(e, 1), d_) for e, d_ in per[s_]))

#!/usr/bin/env python3
"""Write profiles/README.md from the committed bench lines of one capture.  Usage: python scripts/make_profiles_readme.py <tag>"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
P = lambda n: os.path.join(ROOT, "profiles", n)
j = json.load(open(P(f"{tag}_bench.json"))); r = json.load(open(P(f"{tag}_bench_reference.json"))); t = json.load(open(P("traffic.json")))
n2 = None
import glob
def multi(n):
    """Bench line of an N-GPU run: this capture's if present, else the latest earlier one."""
    own = P(f"{tag}_bench_n{n}.json")
    if os.path.isfile(own):
        return os.path.basename(own)
    older = sorted(glob.glob(P(f"r1_?_bench_n{n}.json")))
    return os.path.basename(older[-1]) if older else None
n2file = multi(2)
if n2file:
    n2 = json.load(open(P(n2file)))
n2path = None
hist = [("r1_a (first CUDA path)", "65.6 GDE/s", "0.909", "0.652 ms", "—"),
        ("r1_c (session 2)", "123.5 GDE/s", "0.483", "0.336 ms", "69.1 GDE/s (0.863 ms)"),
        ("r1_e (session 3, mid)", "136.5 GDE/s", "0.437", "0.290 ms", "73.1 GDE/s (0.815 ms)")]
txt = f"""# profiles/ — measured numbers and ncu evidence (round 1)

Every file is named per capture: `r1_a` (first CUDA path), `r1_c` (end of session 2), `r1_e` / `{tag}` (session 3; `{tag}` is current).
All runs: one NVIDIA B200 (148 SMs, SM clock 1965 MHz during the timed regions, no throttle reason), image of this
repository, `scripts/gpu_round.sh <tag>` = smoke → `pytest -m gpu` → `bench.py` → `bench.py --impl reference` →
ncu launch lists (`--metrics gpu__time_duration.sum --clock-control none`; `{tag}_launches_bench.csv` is the list of the
bench command itself, `{tag}_launches.csv` of `profiles/prof_frame.py`) → one `ncu --set full --clock-control none
--import-source on` capture each of the aggregation and the WTA kernel. Numbers printed by a run under ncu are never
used as bench values.

## Current ({tag}) — config C2: 1242×375, D=128, 8 paths, census 5×5, LR + uniqueness, random-texture pair

| quantity | value | file |
|---|---|---|
| `value` (device-resident hot path, census..LR check) | **{j['value']/1e3:.1f} GDE/s = {j['frames_per_s']:.0f} frames/s, {j['ms_per_step']:.3f} ms per frame** | `{tag}_bench.json` |
| single-frame latency, host sync + L2 flush between frames | median {j['latency_ms']['median']:.3f} ms, p95 {j['latency_ms']['p95']:.3f} ms (target < 1 ms) | `{tag}_bench.json` |
| `e2e` (`SGM_Match` with pinned host buffers: H2D + hot path + speckle filter + in-place median + D2H) | **{j['e2e']['value']/1e3:.1f} GDE/s, {j['e2e']['ms_per_step']:.3f} ms per frame** (hot path only through host buffers: {j['e2e']['hotpath_only']['ms_per_step']:.3f} ms) | `{tag}_bench.json` |
| 4 frames in flight, 8 paths | {j['batched']['value']/1e3:.1f} GDE/s = {j['batched']['frames_per_s']:.0f} frames/s | `{tag}_bench.json` `batched` |
| config C4's shape (4 paths), 4 frames in flight | {j['batched_c4']['value']/1e3:.1f} GDE/s = {j['batched_c4']['frames_per_s']:.0f} frames/s per GPU | `{tag}_bench.json` `batched_c4` |
| 9×7 / 64-bit census extension, device-resident hot path | {j['census9x7']['value']/1e3:.1f} GDE/s, {j['census9x7']['ms_per_step']:.3f} ms per frame (aggregation {j['census9x7']['aggregation_kernel_ms']:.3f} ms) | `{tag}_bench.json` `census9x7` |
| reference arm: the reference's own `SemiGlobalMatching.c`, {r['cpu_baseline']['cores']} host cores, `SGM_Initialize`+`SGM_Match` | {r['value']:.0f} MDE/s = {r['frames_per_s']:.2f} frames/s ({r['cpu_baseline']['single_frame_seconds']:.2f} s per frame per core) | `{tag}_bench_reference.json` |
| `cpu_baseline` inside the GPU arm's line (same code, hot-path span only) | {j['cpu_baseline']['value']:.0f} MDE/s on {j['cpu_baseline']['cores']} cores | `{tag}_bench.json` |
"""
if n2:
    txt += f"| 2 GPUs under torchrun (one frame per rank per step, no collective) | {n2['value']/1e3:.1f} GDE/s, {n2['ms_per_step']:.3f} ms per step; C4 shape {n2['batched_c4']['value']/1e3:.1f} GDE/s | `{n2file}` |\n"
if os.path.isfile(P(f"{tag}_configs.jsonl")):
    txt += f"""
## All BASELINE.json config shapes, device-resident hot path, one frame at a time (`{tag}_configs.jsonl`, `profiles/prof_configs.py`)

| config | ms per frame | GDE/s | K2 ms | fraction of the HBM roofline (algorithmic bytes of SURVEY §8d ÷ time ÷ measured peak) |
|---|---|---|---|---|
"""
    for line in open(P(f"{tag}_configs.jsonl")):
        c = json.loads(line)
        txt += f"| {c['config']} | {c['ms_per_frame']:.3f} | {c['MDE_per_s']/1e3:.1f} | {c['aggregation_ms']:.3f} | {c['frame_roofline_frac']:.2f} |\n"
scal = [(1, j, f"{tag}_bench.json")] + [(n, json.load(open(P(multi(n)))), multi(n)) for n in (2, 4, 8) if multi(n)]
if len(scal) > 1:
    txt += f"""
## Scaling over the GPUs of one box (`bench.py` under torchrun, weak scaling: one frame per rank per step, no data-path collective)

| GPUs | `value` GDE/s | ms per step | vs N×(1 GPU) | `e2e` GDE/s | C4 shape batched GDE/s (frames/s) | file |
|---|---|---|---|---|---|---|
"""
    for n, b, fn in scal:
        txt += (f"| {n} | {b['value']/1e3:.1f} | {b['ms_per_step']:.3f} | {b['value']/(n*j['value']):.3f} | {b['e2e']['value']/1e3:.1f} | "
                f"{b['batched_c4']['value']/1e3:.1f} ({b['batched_c4']['frames_per_s']:.0f}) | `{fn}` |\n")
    if any(not fn.startswith(tag) for _, _, fn in scal):
        txt += "\n(Lines from an earlier capture are named by that capture.)\n"
txt += f"""
Ratio `e2e` ÷ reference arm on the same box: {j['e2e']['value']/r['value']:.0f}× (the driver computes its own).

## Kernels of one frame (ncu launch list of the bench command `{tag}_launches_bench.csv`: cold-cache, serialised; event-timed K2 in brackets)

| kernel | grid × block | time under ncu | share of the hot path |
|---|---|---|---|
| K1 `sgm_census<5,5,uint32_t,false>` | (20,47,2) × 256 | ≈ 11 µs | 2 % |
| K2 `sgm_aggregate_paths<4,16,8,8,2,uint32_t,false>` | 562 × 128 | ≈ 345 µs [{j['roofline']['kernel_ms']*1e3:.0f} µs event-timed inside the bench's timed region = {j['roofline']['kernel_share_of_step']*100:.0f} % of the step; 72 % of the ncu list's K1+K2+K3] | 66–72 % |
| K3 `sgm_reduce_wta_lr<8,8,false>` | 375 × 256 | ≈ 124 µs | 26–30 % |
| K4 `speckle_init` / `speckle_merge` / `speckle_count` | | 10 / 22 / 18 µs | (SGM_Match only) |
| K5 `median_prepare` / `median_wavefront` | | 12 / 145 µs | (SGM_Match only) |
| 9×7 census variants: `sgm_census<9,7,…>` / `sgm_aggregate_paths<4,16,4,16,2,u64,false>` | | 19 / 519 µs | |

## Roofline of the dominant kernel (K2), as `bench.py` reports it
* algorithmic bytes per launch (SURVEY §8d model, S read-modify-written per direction): (4·8 − 2)·1242·375·128 = 1 788 480 000 B;
  achieved = that ÷ {j['roofline']['kernel_ms']*1e3:.0f} µs = {j['roofline']['achieved']:.0f} GB/s = **{j['roofline']['frac']:.2f} of the measured HBM copy peak** (6547.5 GB/s, `MEASURED_PEAKS.json`).
* bytes actually moved (ncu `dram__bytes_read.sum + dram__bytes_write.sum`, `{tag}_ncu_full_sgm_aggregate.txt`): {t['aggregate_dram_bytes_per_launch']/1e6:.0f} MB
  per launch ({t['aggregate_dram_read_mb']:.1f} MB read, {t['aggregate_dram_write_mb']:.0f} MB written): the design writes 8 byte planes once instead of
  read-modify-writing a uint16 S, so the traffic is 4× *below* the algorithmic model — no wasted re-reads.
* what binds it instead: ALU pipe {t['alu_pipe_pct']:.0f} %, XU (POPC) {t['xu_pipe_pct']:.0f} %, issue slots {t['issue_active_pct']:.0f} %, L1 data pipe ≈ 63 % of elapsed — none saturated;
  {t['warp_instructions']/1e6:.0f} M warp instructions in dependent chains with ≈ 3.8 warps per scheduler, plus the single-scoreboard prefetch wait (DESIGN.md §3.2).
* K3 (`{tag}_ncu_full_sgm_reduce_wta.txt`): reads 479 MB → HBM floor 73 µs, measured 124 µs (59 % of its HBM roofline); ALU pipe 59 %, issue 61 %.

## History
| capture | `value` | ms per frame | K2 | `e2e` |
|---|---|---|---|---|
"""
for h in hist:
    txt += "| " + " | ".join(h) + " |\n"
txt += f"| {tag} (session 3, final) | {j['value']/1e3:.1f} GDE/s | {j['ms_per_step']:.3f} | {j['roofline']['kernel_ms']:.3f} ms | {j['e2e']['value']/1e3:.1f} GDE/s ({j['e2e']['ms_per_step']:.3f} ms) |\n"
open(P("README.md"), "w").write(txt)
print("wrote profiles/README.md for", tag)

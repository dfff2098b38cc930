#!/usr/bin/env python3
"""Copy the judged artefacts of one GPU round from gpurun_out/<tag>/ into profiles/ (tracked):
bench lines, the ncu launch list, a text summary of every `ncu --set full` capture, and profiles/traffic.json
(dram bytes per launch of the aggregation kernel, read by bench.py for roofline.traffic).
Usage: python scripts/make_profile_summary.py <tag>"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
for name in ("bench.json", "bench_reference.json", "launches.csv", "launches_bench.csv", "gpu.txt"):
    p = os.path.join(src, name)
    if os.path.isfile(p) and os.path.getsize(p):
        shutil.copy(p, os.path.join(dst, f"{tag}_{name}"))

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return [dict(zip(rows[0], r)) for r in rows[2:]]

facts = {}
for f in sorted(os.listdir(src)):
    if not f.endswith(".ncu-rep"):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(src, f)],
                         capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{tag}_ncu_{f[:-8]}.txt"), "w").write(txt)
    for d in raw(os.path.join(src, f)):
        if "sgm_aggregate" in d.get("Kernel Name", ""):
            num = lambda k: float(d[k].replace(",", "")) if d.get(k) else None
            unit_scale = 1e6   # ncu reports Mbyte here
            facts = {
                "source": f"profiles/{tag}_ncu_{f[:-8]}.txt (ncu --set full --clock-control none, one launch at C2)",
                "aggregate_dram_bytes_per_launch": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) * unit_scale,
                "aggregate_dram_read_mb": num("dram__bytes_read.sum"), "aggregate_dram_write_mb": num("dram__bytes_write.sum"),
                "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "l1tex_throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_active"),
                "dram_throughput_pct": num("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                "warp_instructions": num("smsp__inst_executed.sum"), "duration_us_under_ncu": num("gpu__time_duration.sum"),
            }
if facts:
    json.dump(facts, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    print(json.dumps(facts, indent=1))

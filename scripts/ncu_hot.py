#!/usr/bin/env python3
"""Per-instruction view of the hottest execution class of an ncu report (source page): stall samples by reason and the
instructions that collect them.  Usage: python scripts/ncu_hot.py report.ncu-rep [top]"""
import collections, csv, re, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.DictReader(out[1:]))
f = lambda r, k: float(r.get(k) or 0)
tot = sum(f(r, "# Samples") for r in rows); ex = sum(f(r, "Instructions Executed") for r in rows)
print(out[0][:160]); print("samples", int(tot), "warp instructions", int(ex))
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.Counter()
for r in rows:
    for k in reasons: agg[k] += f(r, k)
print("stall mix:", ", ".join(f"{k[6:]}={v / max(tot,1) * 100:.1f}%" for k, v in agg.most_common(9)))
b = collections.defaultdict(lambda: [0, 0, 0.0])
for r in rows:
    n = int(f(r, "Instructions Executed")); b[n][0] += 1; b[n][1] += n; b[n][2] += f(r, "# Samples")
print("execution classes (exec count: static instrs, share of dynamic instrs, share of samples):")
for n, (k, t, s) in sorted(b.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {n:8d}: {k:5d} {t / ex * 100:5.1f}% {s / max(tot,1) * 100:5.1f}%")
for i in sorted(range(len(rows)), key=lambda i: -f(rows[i], "# Samples"))[:top]:
    r = rows[i]
    st = {k[6:]: int(f(r, k)) for k in reasons if f(r, k) > 0}
    print(f"{i:6d} {f(r, '# Samples') / max(tot,1) * 100:4.1f}% exec={int(f(r, 'Instructions Executed')):8d} {r['Source'].strip()[:60]:60s} {st}")

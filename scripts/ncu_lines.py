#!/usr/bin/env python3
"""Attribute executed instructions / stall samples of one kernel to source lines.
Usage: python scripts/ncu_lines.py file.ncu-rep [top]   (the report must have been captured with --import-source on)"""
import collections, csv, io, subprocess, sys

def main():
    path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = "", None
    per = collections.OrderedDict(); text = {}; tot = tots = 0; cur = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
        if hdr is None or len(r) <= iex: continue
        if r[0].strip():
            cur = (fname, int(r[0])); text[cur] = r[1]; continue
        try: ex = int(r[iex]); sm = int(r[isamp])
        except ValueError: continue
        d = per.setdefault(cur, [0, 0]); d[0] += ex; d[1] += sm; tot += ex; tots += sm
    print("total warp instructions", tot, "samples", tots)
    for key, (ex, sm) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{key[0]:>14s}:{key[1]:<4d} {ex / tot * 100:5.1f}% inst {sm / max(tots, 1) * 100:5.1f}% samp  {text.get(key, '').strip()[:100]}")

if __name__ == "__main__":
    main()

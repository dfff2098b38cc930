#!/usr/bin/env python3
"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their instruction mix.
Usage: cuobjdump -sass lib.so > dump.sass; python scripts/sass_loops.py dump.sass <substring of the mangled name> [min_instrs]"""
import collections
import re
import sys

text = open(sys.argv[1]).read().split("Function : ")
key = sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "FMNMX", "VIADDMNMX", "VIMNMX", "VIADD", "ISETP", "SEL", "LEA", "IABS", "MOV", "VIMNMX3", "IMNMX", "LOP", "SGXT", "BMSK", "FSEL", "FSETP", "FADD", "PLOP3", "VABSDIFF", "VABSDIFF4", "CS2R")
FMA = ("IMAD", "FFMA", "FMUL", "IDP", "HFMA2", "HADD2")
XU = ("POPC", "MUFU", "FLO", "BREV", "I2F", "F2I", "I2FP", "F2F")
for fn in text[1:]:
    name = fn.split("\n", 1)[0].strip()
    if key not in name:
        continue
    ins = []
    for line in fn.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(f"== {name}: {len(ins)} instructions")
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in loops:
        n = hi - lo + 1
        if n < minlen:
            continue
        ops = collections.Counter()
        for _, t in ins[lo:hi + 1]:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t2.split()[0]
            ops[op.split(".")[0]] += 1
        pipes = collections.Counter()
        for op, k in ops.items():
            if op in ALU: pipes["alu"] += k
            elif op in FMA: pipes["fma"] += k
            elif op in XU: pipes["xu"] += k
            elif op in ("LDG", "STG", "LDS", "STS", "LDC", "SHFL", "REDUX", "CREDUX", "ATOM", "RED", "LDL", "STL", "ULDC", "LDGSTS", "MATCH", "VOTE", "R2UR"): pipes["mem/mio:" + op] += k
            else: pipes["other:" + op] += k
        print(f"-- loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n} instrs  " + " ".join(f"{k}={v}" for k, v in sorted(pipes.items())))
        print("   " + " ".join(f"{k}:{v}" for k, v in ops.most_common(40)))

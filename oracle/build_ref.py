#!/usr/bin/env python3
"""Build recipes for the parity checkers.  TEST INFRASTRUCTURE ONLY (see oracle/sgm_oracle.h).

Two artefacts, both written under oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built files
travel to the GPU box where /root/reference does not exist):

* ``libsgm_oracle.so``            -- our own C restatement, oracle/sgm_oracle.c ("port").
* ``libsgm_ref_<W>x<H>x<D>[_v].so`` -- the reference's own SemiGlobalMatching.c compiled VERBATIM from
  where it lies under /root/reference, wrapped by oracle/ref_shim.c ("sanitised oracle", SURVEY.md
  section 8c).  The reference fixes its buffer sizes with three unguarded ``#define MAX_*`` lines, so a
  copy of its *header* with those three numbers rewritten (W, H + guard rows, D) is generated into a
  temporary directory for the compile; nothing from the reference is written into the repository.

  Variants (each a mechanical patch of a temporary copy of the .c, applied by exact string match):
    d256  SGM.c:272  ``for (uint8_t f = 0; f < sgm.disp_range; f++)`` -> ``uint16_t f``; the unpatched
          loop never terminates for D >= 256.  Applied automatically when D >= 256.
    p4    SGM.c:217-220  the four diagonal CostAggregate calls removed (num_paths == 4 semantics; the
          reference itself ignores SGMOption.num_paths).

CLI:  python oracle/build_ref.py [--oracle] [WxHxD[:p4] ...]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_DIR = "/root/reference/SemiGlobalMatching/SemiGlobalMatching"
REF_C = os.path.join(REF_DIR, "SemiGlobalMatching.c")
REF_H = os.path.join(REF_DIR, "SemiGlobalMatching.h")

CFLAGS = ["-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-toplevel-reorder", "-w"]


def reference_available() -> bool:
    return os.path.isfile(REF_C) and os.path.isfile(REF_H)


def guard_rows(w: int, h: int) -> int:
    """Rows appended to every reference buffer so both out-of-bounds diagonal visits stay inside
    zero-initialised BSS that nothing else reads (needs rows*W >= 2*(H-1) pixels, landscape)."""
    return 2 + (2 * h + w - 1) // w


def ref_lib_path(w: int, h: int, d: int, variant: str = "") -> str:
    tag = f"{w}x{h}x{d}" + (f"_{variant}" if variant else "")
    return os.path.join(OUT, f"libsgm_ref_{tag}.so")


def oracle_lib_path() -> str:
    return os.path.join(OUT, "libsgm_oracle.so")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.isfile(s))


def build_oracle(force: bool = False) -> str:
    """Compile oracle/sgm_oracle.c -> oracle/_ref/libsgm_oracle.so (works anywhere gcc exists)."""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "sgm_oracle.c")
    out = oracle_lib_path()
    if not force and _newer(out, src, os.path.join(HERE, "sgm_oracle.h")):
        return out
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", out + ".tmp", src, "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(out + ".tmp", out)
    return out


def _patched_header(w: int, h: int, d: int) -> str:
    text = open(REF_H, encoding="utf-8", errors="replace").read()
    for name, val in (("MAX_IMG_WIDTH", w), ("MAX_IMG_HEIGHT", h + guard_rows(w, h)), ("MAX_DISPARITY_RANGE", d)):
        text, n = re.subn(rf"(#define\s+{name}\s+)\d+", rf"\g<1>{val}", text)
        if n != 1:
            raise RuntimeError(f"reference header: expected exactly one '#define {name} <int>'")
    return text


_PATCHES = {
    "d256": [("for (uint8_t f = 0; f < sgm.disp_range; f++)", "for (uint16_t f = 0; f < sgm.disp_range; f++)")],
    "p4": [(f"CostAggregate(sgm.img_left, sgm.cost_init, sgm.cost_aggr, (int8_t){a}, (int8_t){b});",
            f"/* p4: diagonal ({a},{b}) disabled */")
           for a, b in (("1", " 1"), ("-1", "-1"), ("1", " -1"), ("-1", "1"))],
}


def _patched_source(kinds: list[str]) -> str:
    text = open(REF_C, encoding="utf-8", errors="replace").read()
    for kind in kinds:
        for old, new in _PATCHES[kind]:
            # tolerate the reference's irregular spacing inside the call's argument list
            pat = re.escape(old).replace(r"\ ", r"\s*")
            text, n = re.subn(pat, new.replace("\\", "\\\\"), text)
            if n != 1:
                raise RuntimeError(f"patch '{kind}': pattern matched {n} times: {old}")
    return text


def build_ref(w: int, h: int, d: int, variant: str = "", force: bool = False) -> str | None:
    """Compile the reference for one shape.  Returns the .so path; if /root/reference is absent,
    returns the path when a previously built file exists, else None."""
    if variant not in ("", "p4"):
        raise ValueError(variant)
    out = ref_lib_path(w, h, d, variant)
    if not reference_available():
        return out if os.path.isfile(out) else None
    shim = os.path.join(HERE, "ref_shim.c")
    if not force and _newer(out, shim, REF_C, REF_H, os.path.abspath(__file__)):
        return out
    os.makedirs(OUT, exist_ok=True)
    kinds = (["d256"] if d >= 256 else []) + (["p4"] if variant == "p4" else [])
    with tempfile.TemporaryDirectory(prefix="sgm_ref_build_") as tmp:
        with open(os.path.join(tmp, "SemiGlobalMatching.h"), "w") as f:
            f.write(_patched_header(w, h, d))
        src_c = REF_C
        if kinds:
            src_c = os.path.join(tmp, "SemiGlobalMatching_patched.c")
            with open(src_c, "w") as f:
                f.write(_patched_source(kinds))
        big = (w * (h + guard_rows(w, h)) * d * 3) > (1 << 30)
        cmd = (["gcc"] + CFLAGS + (["-mcmodel=large"] if big else []) +
               [f"-I{tmp}", f'-DSGM_REF_C="{src_c}"', "-o", out + ".tmp", shim, "-lm"])
        subprocess.run(cmd, check=True)
    os.replace(out + ".tmp", out)
    return out


DEMO_DIR = os.path.join(OUT, "demo")
DEMO_EXE = os.path.join(DEMO_DIR, "bin", "sgm_demo")
REF_DATA = "/root/reference/SemiGlobalMatching/Data"


def build_demo(force: bool = False) -> str | None:
    """Stage the drop-in acceptance run: the reference's UNMODIFIED demo driver main.c, compiled
    against include/SemiGlobalMatching.h and linked to libsgm_b200.so INSTEAD of the reference's SemiGlobalMatching.c,
    plus the two cone images it loads from ../Data/cone/ (main.c:19-20).  Everything goes to oracle/_ref/demo/ (git-ignored,
    shipped to the GPU box); the rpath is relative so the binary finds the library wherever the tree is unpacked.
    Returns the executable's path, or None when neither the reference tree nor a previously staged copy exists."""
    if not reference_available():
        return DEMO_EXE if os.path.isfile(DEMO_EXE) else None
    root = os.path.dirname(HERE)
    include = os.path.join(root, "include")
    libdir = os.path.join(root, "soc_project_stereo_matching_b200", "lib")
    main_c = os.path.join(REF_DIR, "main.c")
    data_dst = os.path.join(DEMO_DIR, "Data", "cone")
    if not force and _newer(DEMO_EXE, main_c, os.path.join(include, "SemiGlobalMatching.h"), os.path.abspath(__file__)) \
            and os.path.isfile(os.path.join(data_dst, "im6.png")):
        return DEMO_EXE
    os.makedirs(os.path.dirname(DEMO_EXE), exist_ok=True)
    os.makedirs(data_dst, exist_ok=True)
    import shutil
    for name in ("im2.png", "im6.png"):
        shutil.copyfile(os.path.join(REF_DATA, "cone", name), os.path.join(data_dst, name))
    with tempfile.TemporaryDirectory(prefix="sgm_demo_build_") as tmp:
        # main.c includes "SemiGlobalMatching.h" relative to its own directory; a one-line wrapper puts OUR header (same
        # include guard) in front of it, so the demo source itself is compiled byte for byte as it lies in the reference
        wrapper = os.path.join(tmp, "demo.c")
        with open(wrapper, "w") as f:
            f.write(f'#include "{include}/SemiGlobalMatching.h"\n#include "{main_c}"\n')
        cmd = ["gcc", "-O2", "-std=gnu11", "-w", f"-I{REF_DIR}", "-o", DEMO_EXE + ".tmp", wrapper, f"-L{libdir}", "-lsgm_b200",
               "-Wl,-rpath,$ORIGIN/../../../../soc_project_stereo_matching_b200/lib", "-lm"]
        subprocess.run(cmd, check=True)
    os.replace(DEMO_EXE + ".tmp", DEMO_EXE)
    return DEMO_EXE


def main(argv: list[str]) -> int:
    if "--oracle" in argv or len(argv) == 0:
        print(build_oracle(force=True))
    for spec in argv:
        if spec.startswith("--"):
            continue
        shape, _, variant = spec.partition(":")
        w, h, d = (int(x) for x in shape.lower().split("x"))
        print(build_ref(w, h, d, variant, force=True))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

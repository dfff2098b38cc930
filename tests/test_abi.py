"""CPU: the C-ABI boundary.  The library loads without a GPU, exports every function the headers declare,
keeps the reference's struct layout and signatures, and fails (never falls back) when no device exists."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import soc_project_stereo_matching_b200 as sgm
from helpers import gpu_available
from soc_project_stereo_matching_b200 import build as lib_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
REF_DIR = "/root/reference/SemiGlobalMatching/SemiGlobalMatching"


def declared_functions(header: str) -> list[str]:
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b(SGMB?_[A-Za-z0-9]+)\s*\(", text)


def test_library_is_built_in_tree():
    assert os.path.isfile(lib_build.LIB), "run `python -m soc_project_stereo_matching_b200.build`"
    assert os.path.commonpath([lib_build.LIB, ROOT]) == ROOT


def test_every_declared_symbol_is_exported():
    names = set(declared_functions("SemiGlobalMatching.h")) | set(declared_functions("sgm_b200.h"))
    assert {"SGM_Initialize", "SGM_Reset", "SGM_Match", "SGMB_Create", "SGMB_Match", "SGMB_MatchBatchMultiGPU"} <= names
    raw = C.CDLL(lib_build.LIB)
    missing = [n for n in sorted(names) if not hasattr(raw, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_option_struct_layout_matches_reference_abi():
    """SemiGlobalMatching.h:24-40 on x86-64 SysV: sizeof 28, offsets 0/2/4/6/8/12/16/20/22/24/26."""
    want = dict(num_paths=0, min_disparity=2, max_disparity=4, is_check_unique=6, uniqueness_ratio=8, is_check_lr=12,
                lrcheck_thres=16, is_remove_speckles=20, min_speckle_area=22, p1=24, p2_init=26)
    assert C.sizeof(sgm.SGMOption) == 28 and C.alignment(sgm.SGMOption) == 4
    for k, off in want.items():
        assert getattr(sgm.SGMOption, k).offset == off, k
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "SemiGlobalMatching.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(SGMOption), offsetof(SGMOption, num_paths),
               offsetof(SGMOption, min_disparity), offsetof(SGMOption, max_disparity), offsetof(SGMOption, is_check_unique),
               offsetof(SGMOption, uniqueness_ratio), offsetof(SGMOption, is_check_lr), offsetof(SGMOption, lrcheck_thres),
               offsetof(SGMOption, is_remove_speckles), offsetof(SGMOption, min_speckle_area), offsetof(SGMOption, p1),
               offsetof(SGMOption, p2_init));
        return 0; }'''
    outputs = []
    dirs = [INCLUDE] + ([REF_DIR] if os.path.isdir(REF_DIR) else [])
    for inc in dirs:
        with tempfile.TemporaryDirectory() as tmp:
            open(os.path.join(tmp, "t.c"), "w").write(src)
            exe = os.path.join(tmp, "t")
            subprocess.run(["gcc", "-std=gnu11", f"-I{inc}", "-o", exe, os.path.join(tmp, "t.c")], check=True)
            outputs.append(subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert outputs[0] == ["28", "0", "2", "4", "6", "8", "12", "16", "20", "22", "24", "26"]
    assert all(o == outputs[0] for o in outputs), "our header and the reference's header disagree on SGMOption"


def test_headers_compile_as_c_and_cxx():
    for compiler, std, ext in (("gcc", "-std=c11", "c"), ("g++", "-std=c++17", "cpp")):
        with tempfile.TemporaryDirectory() as tmp:
            f = os.path.join(tmp, "t." + ext)
            open(f, "w").write('#include "sgm_b200.h"\nint main(void){ SGMOption o; (void)o; return SGMB_OK; }\n')
            subprocess.run([compiler, std, "-Wall", "-Werror", f"-I{INCLUDE}", "-fsyntax-only", f], check=True)


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not present")
def test_reference_demo_links_against_the_library_unmodified():
    """Drop-in proof: the reference's own main.c compiles against OUR header and links against libsgm_b200.so
    (stb headers come from the reference tree; nothing is copied)."""
    with tempfile.TemporaryDirectory() as tmp:
        # our header must shadow the reference's: main.c includes "SemiGlobalMatching.h" relative to its own directory,
        # so compile a one-line wrapper that includes ours first (same include guard) and then the demo source.
        wrapper = os.path.join(tmp, "demo.c")
        open(wrapper, "w").write(f'#include "{INCLUDE}/SemiGlobalMatching.h"\n#include "{REF_DIR}/main.c"\n')
        exe = os.path.join(tmp, "demo")
        libdir = os.path.dirname(lib_build.LIB)
        subprocess.run(["gcc", "-std=gnu11", "-w", f"-I{REF_DIR}", "-o", exe, wrapper, f"-L{libdir}", "-lsgm_b200",
                        f"-Wl,-rpath,{libdir}", "-lm"], check=True)
        assert os.path.isfile(exe)


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device every compute entry point fails loudly; nothing is computed on the CPU."""
    assert sgm.lib.SGMB_DeviceCount() < 0
    assert sgm.SGM_Initialize(64, 48, sgm.default_option()) is False
    assert "cuda" in sgm.last_error().lower()
    out = np.zeros((48, 64), np.float32)
    img = np.zeros((48, 64), np.uint8)
    assert sgm.SGM_Match(img, img, out) is False
    with pytest.raises(sgm.SGMError):
        sgm.Context(0)


def test_host_side_path_topology_matches_oracle(oracle):
    """The product's path walker (shared by host classification and the aggregation kernel) against the oracle's."""
    from pyoracle import DIRECTIONS
    for w, h in [(20, 12), (16, 16), (9, 8), (8, 9), (12, 20), (450, 375)]:
        for r, (dx, dy) in enumerate(DIRECTIONS):
            npaths = h if r < 2 else w
            step = 1 if (w, h) != (450, 375) else 37
            for i in range(0, npaths, step):
                assert np.array_equal(sgm.debug_walk_path(w, h, r, i), oracle.walk(w, h, dx, dy, i)), (w, h, r, i)
    # landscape: exactly one irregular path per diagonal direction
    for r, want in zip(range(4, 8), (0, 1241, 0, 1241)):
        assert list(np.nonzero(sgm.debug_classify_paths(1242, 375, r))[0]) == [want]
    assert not sgm.debug_classify_paths(1242, 375, 0).any() and not sgm.debug_classify_paths(1242, 375, 2).any()


def _build_example(tmp):
    exe = os.path.join(tmp, "frame_loop")
    libdir = os.path.dirname(lib_build.LIB)
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", f"-I{INCLUDE}", "-o", exe, os.path.join(ROOT, "examples", "frame_loop.c"),
                    f"-L{libdir}", "-lsgm_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    return exe


def test_plain_c_example_compiles_and_links():
    """examples/frame_loop.c: a C caller using only include/*.h (the reference's call sequence + the frame loop)."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = _build_example(tmp)
        if not gpu_available():
            res = subprocess.run([exe], capture_output=True, text=True)
            assert res.returncode != 0 and "SGM_Initialize" in res.stderr      # fails loudly, no CPU fallback


@pytest.mark.gpu
def test_plain_c_example_runs_on_the_gpu():
    with tempfile.TemporaryDirectory() as tmp:
        res = subprocess.run([_build_example(tmp)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        assert "within 0.5 px of the true shift 9" in res.stdout and "reply of" in res.stdout
        valid, total, close = (int(x) for x in re.findall(r"(\d+) of (\d+) pixels valid, (\d+) within", res.stdout)[0])
        assert valid > 0.5 * total and close > 0.99 * valid

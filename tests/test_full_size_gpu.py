"""Full-size parity (BASELINE.json configs C3, C4, C5) against fixtures generated from the COMPILED REFERENCE.

tests/golden/full_c3.npz / full_c5.npz / full_c4.npz were produced by tests/golden/make_golden_full.py from the
reference's own SemiGlobalMatching.c (d256 / p4 one-token patches, guard rows) on `make_pair` inputs.  The volumes are
too large to commit, so the fixtures hold md5 digests of every stage plus row / column CRCs that localise a mismatch.
Everything is compared bit-exactly (tolerance 0, sub-pixel values included)."""
import zlib

import numpy as np
import pytest

import soc_project_stereo_matching_b200 as sgm
from helpers import GOLDEN, md5, to_sgm_option
from pyoracle import options
from soc_project_stereo_matching_b200.synth import make_pair

pytestmark = pytest.mark.gpu

DISP_STAGES = ["disp_left_wta", "disp_right", "disp_lr", "disp_speckle", "disp_final"]


def _row_crc(a: np.ndarray) -> np.ndarray:
    return np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in a], np.uint32)


def _explain(z, tex: str, stage: str, got: np.ndarray) -> str:
    """Localise a mismatch from the stored row / column CRCs, valid count and stored rows."""
    rows = np.nonzero(_row_crc(got) != z[f"{tex}_{stage}_rowcrc"])[0]
    cols = np.nonzero(_row_crc(np.ascontiguousarray(got.T)) != z[f"{tex}_{stage}_colcrc"])[0]
    msg = (f"{tex}/{stage}: md5 differs; {len(rows)} rows (first {rows[:8].tolist()}), {len(cols)} columns (first {cols[:8].tolist()}) "
           f"differ; valid {int(np.isfinite(got).sum())} vs {int(z[f'{tex}_{stage}_valid'])}")
    key = f"{tex}_{stage}_rows"
    if key in z.files:
        step = int(z["row_step"])
        want = z[key]
        bad = np.argwhere(got[::step].view(np.uint32) != want.view(np.uint32))
        if len(bad):
            r, c = int(bad[0][0]), int(bad[0][1])
            msg += f"; first stored-row mismatch at ({r * step},{c}): got {got[r * step, c]!r}, want {want[r, c]!r}"
    return msg


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_full_size_against_compiled_reference(name):
    """C3 (2864x1924, D=256) and C5 (3840x2160, D=256), 8 paths, LR check, sub-pixel: S and every disparity stage of
    SGM_Match's pipeline equal the compiled reference bit for bit, on both synthetic textures."""
    z = np.load(f"{GOLDEN}/full_{name}.npz")
    w, h, d = (int(x) for x in z["shape"])
    opts = options(max_disparity=d)
    errors = []
    with sgm.Context(0) as c:
        c.set_pipeline(sgm.PIPE_REFERENCE | sgm.PIPE_TAPS)
        c.configure(w, h, to_sgm_option(opts))
        for tex in ("noise", "scene"):
            left, right, truth = make_pair(w, h, d, seed=0xB200, texture=tex)
            assert md5(left) == str(z[f"{tex}_md5_left"]) and md5(right) == str(z[f"{tex}_md5_right"]), "input generator drifted"
            got = {"disp_final": c.match(left, right)}
            for k in DISP_STAGES[:-1]:
                got[k] = c.stage(k)
            aggr = c.stage("aggr")
            if md5(aggr) != str(z[f"{tex}_aggr_md5"]):
                rows = np.nonzero(aggr.reshape(h, -1).sum(axis=1, dtype=np.uint64) != z[f"{tex}_aggr_rowsum"])[0]
                errors.append(f"{tex}/aggr: md5 differs; sum {int(aggr.sum(dtype=np.uint64))} vs {int(z[f'{tex}_aggr_sum'])}; "
                              f"{len(rows)} rows differ (first {rows[:8].tolist()})")
            del aggr
            for k in DISP_STAGES:
                if md5(got[k]) != str(z[f"{tex}_{k}_md5"]):
                    errors.append(_explain(z, tex, k, got[k]))
            if tex == "noise":      # sanity of the synthetic input: the known shift is recovered
                a = got["disp_lr"]
                v = np.isfinite(a)
                assert v.mean() > 0.5 and (np.abs(a[v] - truth[v]) <= 0.5).mean() > 0.99
                assert np.all(a[~v].view(np.uint32) == 0x7F800000)
        # the hot-path pipeline (no taps, no post-processing) returns the post-LR map of the last texture
        c.set_pipeline(sgm.PIPE_HOTPATH)
        hot = c.match(left, right)
        if md5(hot) != str(z["scene_disp_lr_md5"]):
            errors.append(_explain(z, "scene", "disp_lr", hot) + " (hot-path pipeline)")
    assert not errors, "\n".join(errors)


def test_c4_all_256_pairs_against_compiled_reference():
    """Config C4 as BASELINE.json states it: 256 KITTI-shaped pairs, D=128, 4 paths, sharded over every visible GPU by
    the library (SGMB_Pool*); the post-LR map of every pair equals the p4-patched compiled reference (md5)."""
    z = np.load(f"{GOLDEN}/full_c4.npz")
    w, h, d, n = (int(x) for x in z["shape"])
    opts = options(max_disparity=d, num_paths=4)
    lefts = np.empty((n, h, w), np.uint8); rights = np.empty((n, h, w), np.uint8)
    for k in range(n):
        lefts[k], rights[k], _ = make_pair(w, h, d, seed=0xB200 + k, texture="scene" if k % 3 == 0 else "noise")
    assert [md5(lefts[k]) for k in (0, 1, n - 1)] == [str(z["md5_left"][k]) for k in (0, 1, n - 1)], "input generator drifted"
    ndev = sgm.lib.SGMB_DeviceCount()
    with sgm.Pool(list(range(ndev)), slots_per_device=4) as pool:
        pool.configure(w, h, to_sgm_option(opts), sgm.PIPE_HOTPATH)
        got = pool.match_batch(lefts, rights)
    bad = [k for k in range(n) if md5(got[k]) != str(z["md5_hotpath"][k])]
    assert not bad, f"{len(bad)} of {n} pairs differ from the reference: {bad[:16]}"
    assert [int(np.isfinite(got[k]).sum()) for k in range(n)] == z["valid"].tolist()

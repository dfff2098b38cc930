"""CPU: the independent numpy restatement (oracle/np_model.py) against the C oracle - census, cost volume, every
direction's path costs and S - for the reference's 5x5 census (which the compiled reference pins) AND the 9x7 / 64-bit
extension (which has no reference code: this is its only independent check beyond the oracle itself)."""
import numpy as np
import pytest

import np_model
from pyoracle import DIRECTIONS, options
from soc_project_stereo_matching_b200.synth import make_pair

CASES = [
    # w, h, texture, census, option overrides
    (48, 24, "scene", (5, 5), dict(max_disparity=32)),
    (48, 24, "scene", (9, 7), dict(max_disparity=32)),
    (64, 32, "noise", (9, 7), dict(max_disparity=64, p1=7, p2_init=90)),
    (64, 32, "scene", (5, 5), dict(max_disparity=64, p1=300, p2_init=3000)),          # penalties beyond uint8
    (40, 30, "scene", (9, 7), dict(max_disparity=24, min_disparity=5, num_paths=4)),
    (33, 21, "scene", (9, 7), dict(max_disparity=40)),                                # D > W: many out-of-row costs
    (20, 20, "noise", (5, 5), dict(max_disparity=8)),                                 # square: wraps on every diagonal
    (14, 22, "scene", (9, 7), dict(max_disparity=8)),                                 # portrait: several irregular paths
]


@pytest.mark.parametrize("w,h,tex,win,kw", CASES)
def test_numpy_model_equals_c_oracle(oracle, w, h, tex, win, kw):
    opts = options(census_w=win[0], census_h=win[1], **kw)
    dmin, dmax = opts["min_disparity"], opts["max_disparity"]
    left, right, _ = make_pair(w, h, dmax - dmin, seed=0x5EED + w, texture=tex)
    if w == 20:
        left[5:12, 3:15] = 200                     # flat block: adaptive P2 at its maximum, uint8 wrap of L_r
    got = oracle.match(left, right, opts, per_direction=True)
    cl, cr = np_model.census(left, *win), np_model.census(right, *win)
    assert np.array_equal(cl, got["census_left"].astype(np.uint64)) and np.array_equal(cr, got["census_right"].astype(np.uint64))
    cost = np_model.cost_volume(cl, cr, dmin, dmax)
    assert np.array_equal(cost, got["cost"]), "cost volume"
    total, per = np_model.aggregate(left, cost, opts["p1"], opts["p2_init"], opts["num_paths"])
    for r, a in enumerate(per):
        assert np.array_equal(a, got["path_cost"][r]), f"direction {DIRECTIONS[r]}"
    assert np.array_equal(total, got["aggr"])


def _random_cases(n=20, seed=97):
    """Seeded draws: both census windows, landscape / square / portrait, D up to 96, penalties up to beyond uint8."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w, h = int(rng.integers(10, 57)), int(rng.integers(8, 41))
        d = int(rng.integers(1, 97))
        mind = int(rng.integers(0, 6)) if i % 3 == 0 else 0
        out.append((w, h, str(rng.choice(["scene", "noise"])), (9, 7) if i % 2 else (5, 5),
                    dict(min_disparity=mind, max_disparity=mind + d, num_paths=int(rng.choice([4, 8])),
                         p1=int(rng.choice([0, 5, 10, 40, 300])), p2_init=int(rng.choice([0, 90, 150, 400, 3000])))))
    return out


@pytest.mark.parametrize("w,h,tex,win,kw", _random_cases())
def test_numpy_model_equals_c_oracle_on_random_draws(oracle, w, h, tex, win, kw):
    test_numpy_model_equals_c_oracle(oracle, w, h, tex, win, kw)


def test_numpy_walker_equals_the_library_and_the_oracle(oracle):
    import soc_project_stereo_matching_b200 as sgm
    for w, h in [(20, 12), (16, 16), (9, 8), (8, 9), (12, 20)]:
        for r, (dx, dy) in enumerate(DIRECTIONS):
            for i in range(h if r < 2 else w):
                want = np_model.walk(w, h, dx, dy, i)
                assert list(oracle.walk(w, h, dx, dy, i)) == want, (w, h, r, i)
                assert list(sgm.debug_walk_path(w, h, r, i)) == want, (w, h, r, i)

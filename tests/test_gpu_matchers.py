"""GPU parity of the SAD / NCC / ASW matchers against the oracle.

SAD costs and disparities are integer work: bit-exact.  NCC (float64 in the reference, closed form from exact integer
window sums here) and ASW (float64 sums in the reference, row-blocked float32 sums here) are floating-point stages
(BASELINE.json north_star): cost-volume relative error <= 1e-5, >= 99.9 % identical disparities, the remainder within +-1
-- or a tie: a pixel whose two candidate disparities have reference costs within the same 1e-5 of each other is decided
by rounding noise below the stated cost tolerance, however far apart the candidates are.
"""
import numpy as np
import pytest

from stereo_match_traditional_b200.synth import make_pair, replicate_pad

pytestmark = pytest.mark.gpu


def cost_rel_err(got, ref, floor=1e-30):
    """max |got - ref| / max(|ref|, floor) over cells where the reference is finite (NaN cells must be NaN on both sides).
    `floor`: NCC correlations that are exactly 0 here (the numerator is an exact integer) come out of the reference's
    double-precision mean-centred sums as ~1e-17 rounding residue; below 1e-9 the error is measured absolutely."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN cells differ"
    ok = ~np.isnan(ref)
    r = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), floor)
    r[got[ok] == ref[ok]] = 0
    return float(r.max()) if r.size else 0.0


def assert_disparities(got, ref, ref_costs, what, frac=0.999):
    """>= frac identical; every other pixel within +-1 or a tie within 1e-5 between the reference costs of the two picks."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    same = got == ref
    assert same.mean() >= frac, f"{what}: {same.mean():.4%} identical"
    ii, jj = np.nonzero(~same)
    for i, j in zip(ii, jj):
        a, b = int(got[i, j]), int(ref[i, j])
        if abs(a - b) <= 1:
            continue
        ca, cb = float(ref_costs[i, j, a]), float(ref_costs[i, j, b])
        assert abs(ca - cb) <= 1e-5 * max(abs(ca), abs(cb)), f"{what}: pixel ({i},{j}) got {a} (cost {ca!r}), reference {b} (cost {cb!r})"

CASES = [(30, 80, 16, "scene", 201), (24, 60, 40, "noise", 202), (20, 50, 8, "constant", 203), (70, 150, 70, "scene", 204),
         (66, 131, 33, "noise", 205)]


@pytest.mark.parametrize("H,W,D,kind,seed", CASES)
@pytest.mark.parametrize("winsize", [3, 1])
def test_sad_bit_exact(engine, oracle, H, W, D, kind, seed, winsize):
    p = make_pair(H, W, D, seed, kind)
    pad = winsize + 1
    lp, rp = replicate_pad(p["gray_left"], pad), replicate_pad(p["gray_right"], pad)
    disp, costs = engine.sad(lp, rp, D, winsize, want_costs=True)
    assert np.array_equal(costs, oracle.sad_costs_left(lp, rp, D, winsize))
    assert np.array_equal(disp, oracle.sad(lp, rp, D, winsize))
    assert np.array_equal(engine.sad(lp, rp, D, winsize, right_view=True), oracle.sad(lp, rp, D, winsize, True))


def test_sad_teddy_shape(engine, oracle):
    """BASELINE.json configs[0]: 9x9 SAD + WTA on a 450x375 pair, 64 disparities."""
    H, W, D = 375, 450, 64
    p = make_pair(H, W, D, 1001, "scene")
    lp, rp = replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)
    got = engine.sad(lp, rp, D, 3)
    assert np.array_equal(got, oracle.sad(lp, rp, D, 3))


@pytest.mark.parametrize("H,W,D,kind,seed", CASES)
@pytest.mark.parametrize("w", [4, 2])
def test_ncc_disparities(engine, oracle, H, W, D, kind, seed, w):
    p = make_pair(H, W, D, seed, kind)
    got, costs = engine.ncc(p["gray_left"], p["gray_right"], D, w, want_costs=True)
    ref = oracle.ncc(p["gray_left"], p["gray_right"], D, w)
    ref_costs = oracle.ncc_costs(p["gray_left"], p["gray_right"], D, w)
    # ComputeCost (NCC.h:15-49) per (pixel, d): 255 where the window leaves the image, NaN for zero-variance windows
    assert cost_rel_err(costs, ref_costs, floor=1e-9) <= 1e-5
    assert_disparities(got, ref, ref_costs, "NCC")
    assert np.all(got[:w] == 0) and np.all(got[:, :w] == 0) and np.all(got[-w:] == 0) and np.all(got[:, -w:] == 0)
    # NCC.h:88: for j < w + D - 1 the first out-of-image disparity (score 255) wins
    j = np.arange(w, min(W - w, w + D - 1))
    if kind != "constant":
        assert np.all(got[w:H - w][:, j] == (j - w + 1)[None, :])


@pytest.mark.parametrize("H,W,D,kind,seed", CASES[:4])
def test_asw_disparities(engine, oracle, H, W, D, kind, seed):
    win = 3
    p = make_pair(H, W, D, seed, kind)
    lp, rp = replicate_pad(p["gray_left"], win + 1), replicate_pad(p["gray_right"], win + 1)
    gl, costs = engine.asw(lp, rp, D, win, 50, 30, 40, want_costs=True)
    rl = oracle.asw(lp, rp, D, win, 50, 30, 40)
    ref_costs = oracle.asw_costs(lp, rp, D, win, 50, 30, 40)
    # costVolume (ASW.h:358-371): float of the double quotient in the reference, row-blocked float32 sums here
    assert cost_rel_err(costs, ref_costs) <= 1e-5
    assert_disparities(gl, rl, ref_costs, "ASW left")
    gr = engine.asw(lp, rp, D, win, 50, 30, 40, right_view=True)
    rr = oracle.asw(lp, rp, D, win, 50, 30, 40, True)
    cut = W - win - 2   # beyond: the reference reads costVolume[-1] (ASW.h:401,:424)
    assert (gr[:, :cut] == rr[:, :cut]).mean() >= 0.999
    assert np.array_equal(engine.asw_crosscheck(rl, rr * 0), oracle.asw_crosscheck(rl, rr * 0))


def test_asw_reference_window(engine, oracle):
    """The shipped parameters (ASWeight.cpp:43-47): winSize 11 -> 25x25 window, sigma 50/30, T 40."""
    H, W, D, win = 30, 90, 20, 11
    p = make_pair(H, W, D, 206, "scene")
    lp, rp = replicate_pad(p["gray_left"], win + 1), replicate_pad(p["gray_right"], win + 1)
    got, costs = engine.asw(lp, rp, D, win, 50, 30, 40, want_costs=True)
    ref = oracle.asw(lp, rp, D, win, 50, 30, 40)
    ref_costs = oracle.asw_costs(lp, rp, D, win, 50, 30, 40)
    assert cost_rel_err(costs, ref_costs) <= 1e-5       # 625 taps per cell
    assert_disparities(got, ref, ref_costs, "ASW 25x25")

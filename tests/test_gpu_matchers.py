"""GPU parity of the SAD / NCC / ASW matchers against the oracle.

SAD costs and disparities are integer work: bit-exact.  NCC (float64 in the reference, closed form from exact integer
window sums here) and ASW (float64 in the reference, float32 accumulation here): >= 99.9 % identical disparities.
"""
import numpy as np
import pytest

from stereo_match_traditional_b200.synth import make_pair, replicate_pad

pytestmark = pytest.mark.gpu

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
    got = engine.ncc(p["gray_left"], p["gray_right"], D, w)
    ref = oracle.ncc(p["gray_left"], p["gray_right"], D, w)
    same = (got == ref).mean()
    assert same >= 0.999, f"{same:.4%} identical"
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
    gl = engine.asw(lp, rp, D, win, 50, 30, 40)
    rl = oracle.asw(lp, rp, D, win, 50, 30, 40)
    assert (gl == rl).mean() >= 0.999, f"left {(gl == rl).mean():.4%}"
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
    got = engine.asw(lp, rp, D, win, 50, 30, 40)
    ref = oracle.asw(lp, rp, D, win, 50, 30, 40)
    assert (got == ref).mean() >= 0.999

"""GPU parity of the AD-Census hot path (cost, aggregation, scanline, post-processing, P_full) against the oracle.

Bars (BASELINE.json north_star): bit-exact for census/Hamming/AD/cost LUT, WTA, LR check; SMT_SUM_REFERENCE_ORDER
aggregation and the scanline are bit-exact too (same operation order); SMT_SUM_FAST is held to rel. err <= 1e-5.
"""
import numpy as np
import pytest

from stereo_match_traditional_b200.synth import make_pair
from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # (H, W, D, kind, seed)
    (48, 200, 24, "scene", 7),
    (33, 257, 40, "scene", 8),
    (64, 64, 16, "noise", 9),
    (40, 96, 128, "scene", 10),   # D > W
    (24, 50, 7, "noise", 11),     # D % 4 != 0 -> scalar paths
    (30, 70, 32, "constant", 12),
]


@pytest.mark.parametrize("H,W,D,kind,seed", SHAPES)
def test_adcensus_cost_bit_exact(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    ref = oracle.adcensus(L, R, D, want=("costL", "costR", "adL", "adR", "censusL", "censusR", "wtaL", "wtaR"))
    for view, s in ((0, "L"), (1, "R")):
        got = engine.adcensus_cost(L, R, D, 10.0, 30.0, view, want_ad=True, want_census=True)
        assert np.array_equal(got["ad"], ref["ad" + s]), f"AD view {view}"
        assert np.array_equal(got["census"], ref["census" + s]), f"Hamming view {view}"
        assert np.array_equal(got["cost"], ref["cost" + s]), f"cost view {view}"
        assert np.array_equal(engine.wta(got["cost"]), ref["wta" + s]), f"WTA view {view}"


def test_census_bitstring_matches_hamming(engine, oracle):
    """popcount(census_u64(L) ^ census_u64(R shifted)) reproduces the reference Hamming volume at interior pixels."""
    H, W, D = 40, 120, 8
    p = make_pair(H, W, D, 21, "noise")
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    cl = engine.census9x7(L)
    cr = engine.census9x7(R)
    ref = oracle.adcensus(L, R, D, want=("censusL",))["censusL"]
    pop = np.vectorize(lambda v: bin(int(v)).count("1"))
    for d in (0, 3, 7):
        x = np.arange(3 + d, W - 3)
        ham = pop(cl[4:H - 4][:, x] ^ cr[4:H - 4][:, x - d])
        assert np.array_equal(ham.astype(np.float32), ref[4:H - 4][:, x, d])


@pytest.mark.parametrize("H,W,D,kind,seed", SHAPES[:5])
def test_cross_aggregate_reference_order_bit_exact(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    rng = np.random.default_rng(seed)
    cost = rng.random((H, W, D), dtype=np.float32) * 2
    ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost)
    got = engine.aggregate_cross(p["bgr_left"], cost, want_arms=True)
    assert np.array_equal(got["arms"], ref["arms"])
    assert np.array_equal(got["sup0"], ref["sup0"])
    assert np.array_equal(got["sup1"], ref["sup1"])
    assert np.array_equal(got["cost"], ref["cost"])


@pytest.mark.parametrize("H,W,D,kind,seed", SHAPES[:5])
def test_scanline_bit_exact(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    rng = np.random.default_rng(seed + 100)
    cost = rng.random((H, W, D), dtype=np.float32) * 2
    gray = p["gray_left"].astype(np.float32)
    names = ("left", "right", "up", "down", "sum", "disp")
    ref = oracle.scanline(cost, gray, 10, 150, want=names)
    got = engine.scanline(cost, gray, 10, 150, want=names)
    for n in names:
        assert np.array_equal(got[n], ref[n]), n
    # fused variant (no path volumes): same disparity
    got2 = engine.scanline(cost, gray, 10, 150, want=("disp",))
    assert np.array_equal(got2["disp"], ref["disp"])


def _disp_pair(H, W, D, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, D, (H, 1)).astype(np.float32) + np.zeros((H, W), np.float32)
    dl = base + rng.integers(-1, 2, (H, W))
    dl[rng.random((H, W)) < 0.08] = rng.integers(0, D)
    dr = base + rng.integers(-1, 2, (H, W))
    dr[rng.random((H, W)) < 0.08] = rng.integers(0, D)
    return np.clip(dl, 0, D - 1).astype(np.float32), np.clip(dr, 0, D - 1).astype(np.float32)


@pytest.mark.parametrize("H,W,D,seed", [(40, 120, 32, 1), (37, 301, 64, 2), (8, 16, 20, 3)])
def test_postprocessing_bit_exact(engine, oracle, H, W, D, seed):
    dl, dr = _disp_pair(H, W, D, seed)
    dl[0, :3] = np.inf  # already-invalid inputs take the :90 branch
    rd, rocc, rmis = oracle.lrcheck(dl, dr, 2)
    gd, gocc, gmis = engine.lrcheck(dl, dr, 2)
    assert np.array_equal(gd, rd)
    assert np.array_equal(gocc, rocc) and np.array_equal(gmis, rmis)
    rl, rocc2, rmis2 = oracle.lrcheck_and(dl[1:], dr[1:], 2.0)
    gl, gocc2, gmis2 = engine.lrcheck_and(dl[1:], dr[1:], 2.0)
    assert np.array_equal(gl, rl)
    assert np.array_equal(gocc2, rocc2) and np.array_equal(gmis2, rmis2)
    # speckles: the shipped call (invalid = +inf -> 2147483647) and a plain integer sentinel
    rs = oracle.remove_speckles(rd, 1, 30, None)
    gs = engine.remove_speckles(gd, 1, 30)
    assert np.array_equal(gs, rs)
    rs2 = oracle.remove_speckles(dl[1:], 1, 12, -1)
    gs2 = engine.remove_speckles(dl[1:], 1, 12, -1)
    assert np.array_equal(gs2, rs2)
    for k in (3, 5):
        assert np.array_equal(engine.median(gs, k), oracle.median(rs, k))


def oracle_pfull(oracle, p, D):
    """P_full from the oracle's stage functions."""
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    c = oracle.adcensus(L, R, D, want=("costL", "costR"))
    aggL = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], c["costL"])["cost"]
    aggR = oracle.cross_aggregate(p["bgr_right"], p["bgr_left"], c["costR"])["cost"]
    sl = oracle.scanline(aggL, L, 10, 150, want=("disp",))["disp"]
    dr = oracle.wta(aggR)
    lrc, _, _ = oracle.lrcheck(sl, dr, 2)
    spk = oracle.remove_speckles(lrc, 1, 30, None)
    return {"left_raw": sl, "right_raw": dr, "lrc": lrc, "disp": oracle.median(spk, 3)}


@pytest.mark.parametrize("H,W,D,kind,seed", [(48, 200, 24, "scene", 31), (60, 90, 32, "noise", 32), (45, 130, 20, "scene", 33)])
def test_pipeline_reference_order_bit_exact(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    ref = oracle_pfull(oracle, p, D)
    got = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"],
                          want=("disp", "left_raw", "right_raw", "lrc"))
    for n in ("right_raw", "left_raw", "lrc", "disp"):
        assert np.array_equal(got[n], ref[n]), n
    # gray computed on the device from BGR (cvtColor fixed-point formula) gives the same answer
    got2 = engine.pipeline(p["bgr_left"], p["bgr_right"], D)
    assert np.array_equal(got2["disp"], ref["disp"])


# ---------------------------------------------------------------------------------------------- SMT_SUM_FAST
FAST_SHAPES = SHAPES[:5] + [(70, 300, 36, "scene", 13), (300, 41, 33, "scene", 14), (12, 700, 8, "scene", 15)]


@pytest.mark.parametrize("H,W,D,kind,seed", FAST_SHAPES)
def test_cross_aggregate_fast_within_tolerance(engine, oracle, H, W, D, kind, seed):
    """Tolerance from BASELINE.json north_star: cost-volume relative error <= 1e-5 (element-wise, max)."""
    from stereo_match_traditional_b200.api import SUM_FAST

    p = make_pair(H, W, D, seed, kind)
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    cost = engine.adcensus_cost(L, R, D)["cost"]          # realistic costs incl. exact zeros
    ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost)["cost"]
    for iters in (1, 2, 4):
        refk = ref if iters == 4 else oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost, iters=iters)["cost"]
        got = engine.aggregate_cross(p["bgr_left"], cost, iters=iters, mode=SUM_FAST)["cost"]
        err, same = rel_err(got, refk)
        assert err <= 1e-5, f"iters={iters}: max rel err {err:.3e} ({same:.2%} bit-identical)"
    # inputs with negative values take the general (TwoSum) prefix instead of the compensated non-negative one
    ncost = cost - np.float32(0.75)
    rn = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], ncost, iters=2)["cost"]
    gn = engine.aggregate_cross(p["bgr_left"], ncost, iters=2, mode=SUM_FAST)["cost"]
    assert np.abs(gn - rn).max() <= 1e-5 * max(1.0, float(np.abs(rn).max()))
    # integer-valued input: the first iteration is exact under any summation order
    icost = np.rint(cost * 100).astype(np.float32)
    r1 = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], icost, iters=1)["cost"]
    g1 = engine.aggregate_cross(p["bgr_left"], icost, iters=1, mode=SUM_FAST)["cost"]
    assert np.array_equal(g1, r1)


@pytest.mark.parametrize("H,W,D,kind,seed", [(48, 200, 24, "scene", 31), (60, 90, 32, "noise", 32), (120, 300, 40, "scene", 34)])
def test_pipeline_fast_disparities(engine, oracle, H, W, D, kind, seed):
    """>= 99.9 % of final disparity pixels identical, the rest within +-1 (north_star)."""
    from stereo_match_traditional_b200.api import SUM_FAST

    p = make_pair(H, W, D, seed, kind)
    ref = oracle_pfull(oracle, p, D)
    got = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"],
                          want=("disp", "left_raw", "right_raw", "lrc"), sum_mode=SUM_FAST)
    for n in ("right_raw", "left_raw"):
        same = (got[n] == ref[n]).mean()
        assert same >= 0.999, f"{n}: {same:.4%}"
        assert np.abs(got[n] - ref[n]).max() <= 1, n
    same = (got["disp"] == ref["disp"]).mean()
    assert same >= 0.999, f"disp: {same:.4%}"


@pytest.mark.parametrize("t1,t2,L1,L2", [(20, 6, 34, 17), (1, 1, 34, 17), (128, 128, 34, 17), (127, 3, 20, 5), (129, 6, 34, 17),
                                         (20, 200, 34, 17), (255, 255, 34, 0), (60, 30, 9, 34)])
def test_cross_arms_threshold_sweep(engine, oracle, t1, t2, L1, L2):
    """BuildArms (cross_aggregator.cpp:135-225) over the whole threshold range: 1..128 runs the packed SIMD walk, larger
    thresholds the scalar one; high-contrast noise makes every rule (tau1 vs centre, vs predecessor, tau2 beyond L2) fire."""
    H, W = 48, 90
    rng = np.random.default_rng(t1 * 7 + t2)
    base = rng.integers(0, 256, (H, W, 3)).astype(np.int32)
    smooth = (np.add.outer(np.arange(H), np.arange(W))[..., None] * np.array([1, 2, 3]) % 256).astype(np.int32)
    mix = np.where(rng.random((H, W, 1)) < 0.15, base, smooth + rng.integers(-3, 4, (H, W, 3)))
    bgr = np.clip(mix, 0, 255).astype(np.uint8)
    cost = np.zeros((H, W, 1), np.float32)
    ref = oracle.cross_aggregate(bgr, bgr, cost, L1=L1, L2=L2, t1=t1, t2=t2, iters=1)
    got = engine.cross_arms(bgr, L1, L2, t1, t2)
    assert np.array_equal(got["arms"], ref["arms"])
    assert np.array_equal(got["sup0"], ref["sup0"]) and np.array_equal(got["sup1"], ref["sup1"])
    assert ref["arms"].max() > 0


DEGENERATE = [(9, 9, 1), (10, 12, 3), (7, 20, 4), (12, 300, 5), (200, 12, 9), (3, 40, 6), (20, 8, 12), (16, 33, 33), (35, 36, 65),
              (1, 50, 4), (50, 1, 1), (2, 2, 2), (71, 73, 31), (36, 140, 97)]


@pytest.mark.parametrize("H,W,D", DEGENERATE, ids=lambda v: str(v))
def test_pipeline_degenerate_shapes(engine, oracle, H, W, D):
    """Images smaller than the census window, than one arm, than one 32-position chunk, D = 1, D > W, D not a multiple of
    32: the reference-order pipeline stays bit-exact, the fast one within the disparity bar."""
    from stereo_match_traditional_b200.api import SUM_FAST

    p = make_pair(H, W, D, 100 + H + W, "noise")
    ref = oracle_pfull(oracle, p, D)
    names = ("disp", "left_raw", "right_raw", "lrc")
    got = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"], want=names)
    for n in names:
        assert np.array_equal(got[n], ref[n], equal_nan=True), n
    fast = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"], want=names, sum_mode=SUM_FAST)
    bar = 0.999 if H * W >= 1000 else 0.99   # below 1000 pixels one differing pixel is already more than 0.1 %
    for n in ("left_raw", "right_raw"):
        assert (fast[n] == ref[n]).mean() >= bar and np.abs(fast[n] - ref[n]).max() <= 1, n


def test_pipeline_caller_owned_output(engine, oracle):
    """`out=`: the disparity map lands in the caller's array (the reference's caller-owned output convention)."""
    H, W, D = 40, 112, 16
    p = make_pair(H, W, D, 5, "scene")
    ref = oracle_pfull(oracle, p, D)
    buf = np.full((H, W), -7.0, np.float32)
    got = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"], out=buf)
    assert got["disp"] is buf and np.array_equal(buf, ref["disp"], equal_nan=True)
    with pytest.raises(ValueError):
        engine.pipeline(p["bgr_left"], p["bgr_right"], D, out=np.zeros((H, W + 1), np.float32))

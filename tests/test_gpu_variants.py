"""GPU: the "remaining variants" of SURVEY.md 8(f)4 and the CBLSM driver sequence, against the compiled reference.

costAggregationVertical / costAggregationV4 (CBLSM/CBLSM.h:1227, :1128), CBLSM's main() call order (CBLSM.cpp:64-153),
the CV_32S CrossCheckDiaparity (SAD/Sad.h:184), the sub-pixel value OptimalDisparity computes and discards (Sad.h:76-84;
checked against the port's restatement -- the reference never returns it) and the shifted-image ncc() (NCC/NCC.h:117)."""
import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST
from stereo_match_traditional_b200.synth import make_pair, replicate_pad

pytestmark = pytest.mark.gpu


def _arms(engine, img):
    a = engine.arms_gray(img, 25, 34, 17, variant=1)
    return {k: a[k] for k in ("left", "right", "up", "down")}


@pytest.mark.parametrize("H,W,D,seed", [(40, 120, 12, 401), (70, 64, 9, 402)])
def test_cblsm_vertical_and_v4_aggregation(engine, oracle, H, W, D, seed):
    p = make_pair(H, W, D, seed, "scene")
    ad = engine.ad_cost(p["gray_left"], p["gray_right"], D, 0)
    arms = _arms(engine, p["gray_left"])
    got = engine.aggregate_rect(ad, arms["left"], arms["right"], arms["up"], arms["down"], variant=0)
    assert np.array_equal(got, oracle.cblsm_aggregate_variant(ad, arms, "vertical"))
    # costAggregationV4: arm VOLUMES, exclusive bounds; empty windows (0/0 = NaN) included
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    av = {"left": np.minimum(rng.integers(0, 6, (H, W, D)), xx[..., None]).astype(np.int32),
          "right": np.minimum(rng.integers(0, 6, (H, W, D)), (W - xx)[..., None]).astype(np.int32),
          "up": np.minimum(rng.integers(0, 6, (H, W, D)), yy[..., None]).astype(np.int32),
          "down": np.minimum(rng.integers(0, 6, (H, W, D)), (H - yy)[..., None]).astype(np.int32)}
    got4 = engine.aggregate_rect(ad, av["left"], av["right"], av["up"], av["down"], variant=3)
    ref4 = oracle.cblsm_aggregate_variant(ad, av, "v4")
    assert np.isnan(ref4).any() and np.array_equal(got4, ref4, equal_nan=True)


@pytest.mark.parametrize("H,W,D,seed", [(60, 150, 20, 411), (90, 100, 32, 412)])
def test_cblsm_pipeline_call_order(engine, oracle, H, W, D, seed):
    """CBLSM.cpp:146-153: R1 = V5(adR, armsR), L1 = V5(adL, armsL), L2 = V5(L1, armsL), R2 = V5(R1, armsL) (left arms, kept)."""
    p = make_pair(H, W, D, seed, "scene")
    gl, gr = p["gray_left"], p["gray_right"]
    aL, aR = oracle.cblsm_arms(gl, 25, 34, 17), oracle.cblsm_arms(gr, 25, 34, 17)
    adL, adR = oracle.cblsm_ad(gl, gr, D)
    L2 = oracle.cblsm_aggregate_v5(oracle.cblsm_aggregate_v5(adL, aL), aL)
    R2 = oracle.cblsm_aggregate_v5(oracle.cblsm_aggregate_v5(adR, aR), aL)
    dl, dr = engine.pipeline_cblsm(gl, gr, D)
    assert np.array_equal(dl, oracle.cblsm_wta(L2)) and np.array_equal(dr, oracle.cblsm_wta(R2))
    fl, fr = engine.pipeline_cblsm(gl, gr, D, mode=SUM_FAST)
    assert (fl == dl).mean() >= 0.999 and (fr == dr).mean() >= 0.999
    assert np.abs(fl - dl).max() <= 1 and np.abs(fr - dr).max() <= 1


def test_sad_int_crosscheck_and_subpixel(engine, oracle, oracle_port):
    H, W, D = 50, 140, 24
    p = make_pair(H, W, D, 421, "scene")
    lp, rp = replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)
    res = engine.sad(lp, rp, D, 3, want_costs=True, want_subpixel=True)
    dl, dr = res["disp"], engine.sad(lp, rp, D, 3, right_view=True)
    assert np.array_equal(dl, oracle.sad(lp, rp, D, 3))
    # the real maps, then maps with planted disagreements of both signs
    rng = np.random.default_rng(5)
    dl2 = dl.copy()
    pick = rng.random(dl.shape) < 0.1
    dl2[pick] = rng.integers(0, 12, int(pick.sum()))
    dl2[0] = np.minimum(dl2[0], np.arange(W))      # row 0: the reference would read before the buffer for lv > j
    for a, b in ((dl, dr), (dl2, dr)):
        out, occ, mis = engine.sad_crosscheck(a, b)
        ro, rocc, rmis = oracle.sad_crosscheck(a, b)
        assert np.array_equal(out, ro) and np.array_equal(occ, rocc) and np.array_equal(mis, rmis)
    assert len(rocc) > 0 and len(rmis) > 0
    # sub-pixel: the restatement of Sad.h:76-84 applied to the reference's own SAD vectors
    costs = oracle.sad_costs_left(lp, rp, D, 3)
    want = np.array([[oracle_port.sad_subpixel(costs[i, j]) for j in range(W)] for i in range(H)], np.float32)
    assert np.array_equal(res["subpixel"], want)
    assert np.array_equal((want != 0), (dl != 0)) and np.abs(want[dl != 0] - dl[dl != 0]).max() <= 0.5 + 1e-6


@pytest.mark.parametrize("right_view,add_constant", [(False, False), (True, True)])
def test_ncc_shift_variant(engine, oracle, right_view, add_constant):
    H, W = 36, 150
    p = make_pair(H, W, 16, 431, "scene")
    got = engine.ncc_shift(p["gray_left"], p["gray_right"], right_view, add_constant)
    ref = oracle.ncc_shift(p["gray_left"], p["gray_right"], right_view, add_constant)
    assert (got == ref).mean() >= 0.999, f"{(got == ref).mean():.4%}"
    assert set(np.unique(got)) <= set(range(0, 240, 3))

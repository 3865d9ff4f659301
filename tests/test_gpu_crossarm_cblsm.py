"""GPU parity of the CrossArmAggregation (AD-CensusV1/CrossArm.cpp) and CBLSM (CBLSM/CBLSM.h) paths against the oracle.

Arm maps, AD volumes and reference-order rectangle means are bit-exact; the summed-area-table mode is held to 1e-5, and
is bit-exact on integer-valued input (first CBLSM pass).
"""
import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST, SUM_REFERENCE_ORDER, SmtError
from stereo_match_traditional_b200.synth import make_pair
from conftest import rel_err

pytestmark = pytest.mark.gpu

# shapes for which CrossArm.cpp:92's flat index stays inside the slice: floor((H*H-1)/W) + 34 < H-1
SAFE = [(48, 200, 6, "scene", 7), (60, 260, 5, "scene", 5), (50, 220, 4, "noise", 6)]
ANY = SAFE + [(64, 64, 4, "scene", 9), (40, 150, 5, "noise", 8), (33, 257, 3, "constant", 4)]


@pytest.mark.parametrize("H,W,D,kind,seed", ANY)
@pytest.mark.parametrize("color", [False, True])
def test_crossarm_arm_maps_bit_exact(engine, oracle, H, W, D, kind, seed, color):
    p = make_pair(H, W, D, seed, kind)
    img = p["bgr_left"] if color else p["gray_left"]
    ref = oracle.crossarm(img, D, 30, 0)
    got = engine.arms_gray(img, 30, 34, 17, variant=0)
    for k in ("left", "right", "top", "bottom"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["tao_final"] == ref["tao_final"]   # 6 once the sticky threshold (CrossArm.cpp:170) has triggered
    cref = oracle.cblsm_arms(img, 25, 34, 17)
    cgot = engine.arms_gray(img, 25, 34, 17, variant=1)
    for k in ("left", "right", "up", "down"):
        assert np.array_equal(cgot[k], cref[k]), k


def test_crossarm_rejects_tall_images(engine):
    """rows > cols: the reference's Right pass (col = _row) indexes outside the image -> unsupported, not garbage."""
    img = np.zeros((40, 30), np.uint8)
    with pytest.raises(SmtError) as e:
        engine.arms_gray(img, 30, variant=0)
    assert e.value.status == 2


@pytest.mark.parametrize("H,W,D,kind,seed", SAFE)
def test_crossarm_aggregation(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    rng = np.random.default_rng(seed)
    vol = rng.random((H, W, D), dtype=np.float32) * 2
    for variant, mode in ((0, 1), (1, 2)):   # AggregationVertical, Aggregation
        ref = oracle.crossarm(p["gray_left"], D, 30, mode, vol)
        arms = [ref[k] for k in ("left", "right", "top", "bottom")]
        got = engine.aggregate_rect(vol, *arms, variant=variant, mode=SUM_REFERENCE_ORDER)
        assert np.array_equal(got, ref["vol"], equal_nan=True), f"variant {variant} reference order"
        fast = engine.aggregate_rect(vol, *arms, variant=variant, mode=SUM_FAST)
        ok = ~np.isnan(ref["vol"])
        assert np.array_equal(np.isnan(fast), ~ok)
        err, _ = rel_err(fast[ok], ref["vol"][ok])
        assert err <= 1e-5, f"variant {variant} fast: {err:.3e}"


def test_flat_index_out_of_bounds_is_refused(engine, oracle):
    """The scrambled Right map can push the flat index past the slice; the engine checks the actual
    addresses and refuses instead of reading out of bounds like the reference."""
    H, W, D = 12, 20, 2
    z = np.zeros((H, W), np.int32)
    right = z.copy()
    right[H - 1, W - 1] = 3          # (H-1)*W + (W-1) + 3 >= H*W
    vol = np.ones((H, W, D), np.float32)
    for variant in (0, 2):
        with pytest.raises(SmtError) as e:
            engine.aggregate_rect(vol, z, right, z, z, variant=variant)
        assert e.value.status == 2
    # the same window one row up only wraps into the next row: allowed for the flat variants, refused for V5
    right = z.copy()
    right[H - 2, W - 1] = 3
    out = engine.aggregate_rect(vol, z, right, z, z, variant=0)
    assert np.all(out == 1)
    with pytest.raises(SmtError):
        engine.aggregate_rect(vol, z, right, z, z, variant=2)


@pytest.mark.parametrize("H,W,D,kind,seed", ANY)
def test_cblsm_pipeline(engine, oracle, H, W, D, kind, seed):
    """CBLSM.cpp:64-67,101-104,133-134,146-153: arms, AD, V5 x2 per view (4th call: LEFT arms on the right volume), WTA."""
    p = make_pair(H, W, D, seed, kind)
    gl, gr = p["gray_left"], p["gray_right"]
    adL_ref, adR_ref = oracle.cblsm_ad(gl, gr, D)
    adL, adR = engine.ad_cost(gl, gr, D, 0), engine.ad_cost(gl, gr, D, 1)
    assert np.array_equal(adL, adL_ref) and np.array_equal(adR, adR_ref)
    aL = oracle.cblsm_arms(gl, 25, 34, 17)
    aR = oracle.cblsm_arms(gr, 25, 34, 17)
    order = ("left", "right", "up", "down")
    R1 = oracle.cblsm_aggregate_v5(adR_ref, aR)
    L1 = oracle.cblsm_aggregate_v5(adL_ref, aL)
    L2 = oracle.cblsm_aggregate_v5(L1, aL)
    R2 = oracle.cblsm_aggregate_v5(R1, aL)
    for mode in (SUM_REFERENCE_ORDER, SUM_FAST):
        gR1 = engine.aggregate_rect(adR, *[aR[k] for k in order], variant=2, mode=mode)
        gL1 = engine.aggregate_rect(adL, *[aL[k] for k in order], variant=2, mode=mode)
        assert np.array_equal(gR1, R1) and np.array_equal(gL1, L1)      # integer input: exact under any order
        gL2 = engine.aggregate_rect(gL1, *[aL[k] for k in order], variant=2, mode=mode)
        gR2 = engine.aggregate_rect(gR1, *[aL[k] for k in order], variant=2, mode=mode)
        if mode == SUM_REFERENCE_ORDER:
            assert np.array_equal(gL2, L2) and np.array_equal(gR2, R2)
        else:
            assert rel_err(gL2, L2)[0] <= 1e-5 and rel_err(gR2, R2)[0] <= 1e-5
        dl, dr = engine.wta(gL2), engine.wta(gR2)
        rl, rr = oracle.cblsm_wta(L2), oracle.cblsm_wta(R2)
        if mode == SUM_REFERENCE_ORDER:
            assert np.array_equal(dl, rl) and np.array_equal(dr, rr)
        else:
            assert (dl == rl).mean() >= 0.999 and (dr == rr).mean() >= 0.999

"""GPU: iterative region voting, proper interpolation and depth-discontinuity adjustment against their fixed specification.

PARITY UNPINNED -- the reference has no implementation of these steps (only the unused option fields
CBLSM/adcensus_types.h:58-59,65); the checker is oracle/port/oracle_port.c (orc_refine_*), which states the specification
(after Mei et al. 2011).  The kernels must match it bit for bit, fed by the real front end: P_full up to the left-right
check on the GPU, then the three steps on both sides."""
import numpy as np
import pytest

from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def _front_end(engine, H, W, D, seed, kind="scene"):
    p = make_pair(H, W, D, seed, kind)
    L, R = p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32)
    aggL = engine.aggregate_cross(p["bgr_left"], engine.adcensus_cost(L, R, D, 10.0, 30.0, 0)["cost"], want_arms=True)
    aggR = engine.aggregate_cross(p["bgr_right"], engine.adcensus_cost(L, R, D, 10.0, 30.0, 1)["cost"])
    dl = engine.scanline(aggL["cost"], L, 10, 150, want=("disp",))["disp"]
    dr = engine.wta(aggR["cost"])
    lrc, occ, mis = engine.lrcheck(dl, dr, 2)
    return p, aggL, lrc, occ, mis


@pytest.mark.parametrize("H,W,D,seed,kind", [(60, 160, 24, 301, "scene"), (90, 120, 40, 302, "scene"), (48, 70, 16, 303, "noise")])
def test_refinement_steps_match_the_fixed_specification(engine, oracle_port, H, W, D, seed, kind):
    p, aggL, lrc, occ, mis = _front_end(engine, H, W, D, seed, kind)
    cls = engine.invalid_class_map(occ, mis, (H, W))
    want_cls = np.zeros((H, W), np.uint8)
    want_cls[occ[:, 0], occ[:, 1]] = 1
    want_cls[mis[:, 0], mis[:, 1]] = 2
    assert np.array_equal(cls, want_cls) and np.array_equal(np.isinf(lrc), cls != 0)
    assert (cls != 0).sum() > 20, "the case must have something to refine"
    # 1. iterative region voting
    d1, c1 = engine.refine_region_voting(lrc, cls, aggL["arms"], D, 20, 0.4, 5)
    r1, k1 = oracle_port.refine_region_voting(lrc, cls, aggL["arms"], D, 20, 0.4, 5)
    assert np.array_equal(d1, r1) and np.array_equal(c1, k1)
    # small regions need a lower ts to vote at all: exercise the accept path too
    d1b, c1b = engine.refine_region_voting(lrc, cls, aggL["arms"], D, 2, 0.3, 3)
    r1b, k1b = oracle_port.refine_region_voting(lrc, cls, aggL["arms"], D, 2, 0.3, 3)
    assert np.array_equal(d1b, r1b) and np.array_equal(c1b, k1b)
    assert (c1b != 0).sum() <= (cls != 0).sum()
    # 2. proper interpolation on what is left
    d2, c2 = engine.refine_interpolate(d1b, c1b, p["bgr_left"], D)
    r2, k2 = oracle_port.refine_interpolate(r1b, k1b, p["bgr_left"], D)
    assert np.array_equal(d2, r2) and np.array_equal(c2, k2)
    assert (c2 != 0).sum() < (c1b != 0).sum() or (c1b != 0).sum() == 0
    assert np.array_equal(np.isinf(d2), c2 != 0)
    # 3. depth-discontinuity adjustment with the aggregated left volume
    d3 = engine.refine_discontinuity(d2, aggL["cost"], 5.0)
    r3 = oracle_port.refine_discontinuity(r2, aggL["cost"], 5.0)
    assert np.array_equal(d3, r3, equal_nan=True)


def test_discontinuity_adjustment_moves_edge_pixels(engine, oracle_port):
    """A crafted map: a vertical disparity step whose edge column is rated better at the neighbour's disparity by the cost
    volume, plus invalid pixels and out-of-range disparities next to the edge."""
    H, W, D = 24, 40, 16
    rng = np.random.default_rng(7)
    disp = np.full((H, W), 3.0, np.float32)
    disp[:, 20:] = 11.0
    disp[5, 19] = np.inf
    disp[9, 21] = 40.0          # beyond D: ignored as a candidate
    cost = rng.random((H, W, D), dtype=np.float32) + 1.0
    cost[:, 19, 11] = 0.1       # the column left of the step prefers the right side's disparity
    cost[:, 20, 3] = 0.2        # and the column right of it the left side's
    got = engine.refine_discontinuity(disp, cost, 5.0)
    ref = oracle_port.refine_discontinuity(disp, cost, 5.0)
    assert np.array_equal(got, ref, equal_nan=True)
    assert (got[1:-1, 19] == 11.0).sum() >= H - 6 and (got[1:-1, 20] == 3.0).sum() >= H - 6
    assert np.array_equal(got[:, :18], disp[:, :18]) and np.isinf(got[5, 19])


def test_refinement_device_tensors_and_errors(engine, oracle_port):
    import torch
    from stereo_match_traditional_b200.api import SmtError

    H, W, D = 50, 90, 20
    p, aggL, lrc, occ, mis = _front_end(engine, H, W, D, 311)
    cls = engine.invalid_class_map(occ, mis, (H, W))
    dev = "cuda:0"
    d1, c1 = engine.refine_region_voting(torch.as_tensor(lrc, device=dev), torch.as_tensor(cls, device=dev),
                                         torch.as_tensor(aggL["arms"], device=dev), D, 2, 0.3, 2)
    r1, k1 = oracle_port.refine_region_voting(lrc, cls, aggL["arms"], D, 2, 0.3, 2)
    assert np.array_equal(d1.cpu().numpy(), r1) and np.array_equal(c1.cpu().numpy(), k1)
    with pytest.raises(SmtError):
        engine.invalid_class_map(np.array([[H, 0]], np.int32), np.zeros((0, 2), np.int32), (H, W))


def test_pipeline_with_refinement_equals_the_composition(engine, oracle_port):
    """smt_pipeline_adcensus_full with the refinement switches on == LR check -> voting -> interpolation -> adjustment ->
    RemoveSpeckles -> MedianFilter composed by hand from the port's steps (and is off by default: P_full unchanged)."""
    H, W, D = 70, 150, 24
    p, aggL, lrc, occ, mis = _front_end(engine, H, W, D, 321)
    got = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"], want=("disp", "lrc"), irv_iters=3, irv_ts=2,
                          irv_th=0.3, do_interpolation=1, do_discontinuity_adjustment=1)
    assert np.array_equal(got["lrc"], lrc, equal_nan=True)
    cls = np.zeros((H, W), np.uint8)
    cls[occ[:, 0], occ[:, 1]] = 1
    cls[mis[:, 0], mis[:, 1]] = 2
    d, c = oracle_port.refine_region_voting(lrc, cls, aggL["arms"], D, 2, 0.3, 3)
    d, c = oracle_port.refine_interpolate(d, c, p["bgr_left"], D)
    d = oracle_port.refine_discontinuity(d, aggL["cost"], 5.0)
    want = oracle_port.median(oracle_port.remove_speckles(d, 1, 30, None), 3)
    assert np.array_equal(got["disp"], want, equal_nan=True)
    plain = engine.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"])
    assert np.isfinite(got["disp"]).sum() > np.isfinite(plain["disp"][plain["disp"] < 1e9]).sum() - 1

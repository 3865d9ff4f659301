"""GPU: the two routes of SMT_SUM_FAST cross aggregation (fixed-point prefixes / float-pair prefixes) against the oracle.

The fixed-point kernels (csrc/cross_agg_fx.cu) carry an error bound that only holds while the values entering a window
are not tiny against the largest cost; lines that hold smaller values are redone by the float-pair kernels queued behind
them, which take over whole sweeps when too many lines ask.  Every mix of the two routes is held to the north_star bar:
cost-volume relative error <= 1e-5 against CrossAggregator::Aggregate (CBLSM/cross_aggregator.cpp:89-118).
"""
import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST
from stereo_match_traditional_b200.synth import make_pair
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _cost(engine, p, D):
    return engine.adcensus_cost(p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32), D)["cost"]


@pytest.mark.parametrize("H,W,D,kind,seed", [(96, 260, 64, "scene", 41), (260, 96, 40, "scene", 42), (80, 120, 33, "noise", 43),
                                             (150, 150, 32, "scene", 44)])
def test_fixed_point_route_is_taken_and_within_tolerance(engine, oracle, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    cost = _cost(engine, p, D)
    for iters in (1, 3, 4):
        ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost, iters=iters)["cost"]
        got = engine.aggregate_cross(p["bgr_left"], cost, iters=iters, mode=SUM_FAST)["cost"]
        redone = engine.aggregate_fast_status()
        ntasks = (H + W) * ((D + 31) // 32) * iters          # (line, chunk) tasks over all passes
        assert 0 <= redone < 0.15 * ntasks, f"AD-Census costs must stay on the fixed-point route ({redone} of {ntasks} redone)"
        err, same = rel_err(got, ref)
        assert err <= 1e-5, f"iters={iters}: max rel err {err:.3e} ({same:.2%} bit-identical)"


@pytest.mark.parametrize("H,W,D,seed", [(64, 200, 32, 51), (200, 64, 24, 52)])
def test_tiny_window_sums_fall_back_to_float_pairs(engine, oracle, H, W, D, seed):
    """2 % of the costs three orders of magnitude below the rest, spread over every line: below the domain of the
    fixed-point bound, so the float-pair passes take the sweeps over, and the result must still meet the bar."""
    p = make_pair(H, W, D, seed, "noise")          # arms ~0: single-tap windows, the worst case for a fixed quantum
    rng = np.random.default_rng(seed)
    cost = _cost(engine, p, D)
    tiny = rng.random(cost.shape) < 0.02
    cost = np.where(tiny, cost * np.float32(1e-3), cost).astype(np.float32)
    ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost)["cost"]
    got = engine.aggregate_cross(p["bgr_left"], cost, mode=SUM_FAST)["cost"]
    assert engine.aggregate_fast_status() == -1, "tiny values on every line must hand the sweeps to the float-pair passes"
    err, same = rel_err(got, ref)
    assert err <= 1e-5, f"max rel err {err:.3e} ({same:.2%} bit-identical)"
    # a few tiny values on a few lines only: those lines are redone, the rest stays fixed-point
    cost = _cost(engine, p, D)
    cost[H // 2, ::7, D // 3] *= np.float32(1e-3)
    ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost)["cost"]
    got = engine.aggregate_cross(p["bgr_left"], cost, mode=SUM_FAST)["cost"]
    assert engine.aggregate_fast_status() > 0
    assert rel_err(got, ref)[0] <= 1e-5
    # and the state is per aggregation: a well-scaled volume afterwards runs fixed-point again
    p2 = make_pair(H, W, D, seed + 1, "scene")
    c2 = _cost(engine, p2, D)
    got2 = engine.aggregate_cross(p2["bgr_left"], c2, mode=SUM_FAST)["cost"]
    assert engine.aggregate_fast_status() >= 0
    assert rel_err(got2, oracle.cross_aggregate(p2["bgr_left"], p2["bgr_right"], c2)["cost"])[0] <= 1e-5


def test_integer_costs_first_iteration_exact_on_both_routes(engine, oracle):
    """CBLSM ComputeAD volumes are integers <= 255: iteration 1 is exact under any summation order."""
    H, W, D = 60, 180, 32
    p = make_pair(H, W, D, 61, "scene")
    ad = engine.ad_cost(p["gray_left"], p["gray_right"], D)
    ref1 = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], ad, iters=1)["cost"]
    got1 = engine.aggregate_cross(p["bgr_left"], ad, iters=1, mode=SUM_FAST)["cost"]
    assert np.array_equal(got1, ref1)
    ref4 = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], ad, iters=4)["cost"]
    got4 = engine.aggregate_cross(p["bgr_left"], ad, iters=4, mode=SUM_FAST)["cost"]
    assert rel_err(got4, ref4)[0] <= 1e-5


def test_fast_sweeps_refuse_long_arms(engine):
    """The streaming kernels keep 72 prefix entries per line; the per-sweep ABI takes caller-built arm maps, so it has to
    look at them (smt_aggregate_cross itself refuses cross_L1 > 34)."""
    import ctypes as C
    import torch

    H, W, D = 40, 200, 32
    p = make_pair(H, W, D, 71, "constant")      # constant image: arms grow to L1
    arms = engine.cross_arms(p["bgr_left"], 60, 17, 20, 6)
    assert arms["arms"].max() > 34
    dev = torch.device("cuda", 0)
    vin = torch.rand((H, W, D), device=dev)
    vout = torch.empty_like(vin)
    a = torch.from_numpy(arms["arms"]).to(dev)
    s0 = torch.from_numpy(arms["sup0"].view(np.int16)).to(dev)
    s1 = torch.from_numpy(arms["sup1"].view(np.int16)).to(dev)
    torch.cuda.synchronize()
    st = engine.lib.smt_aggregate_cross_sweep(engine.h, 4, 0, C.c_void_p(vin.data_ptr()), C.c_void_p(vout.data_ptr()),
                                              C.c_void_p(a.data_ptr()), C.c_void_p(s0.data_ptr()), C.c_void_p(s1.data_ptr()),
                                              H, W, D, 2.0, int(SUM_FAST), 1)
    assert st != 0 and b"arms <= 34" in engine.lib.smt_last_error_string(engine.h)

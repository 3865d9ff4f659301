"""GPU: the row-band split (BASELINE.json configs[4]) emulated on ONE GPU -- every band is a separate buffer with its own
34-row halos, halos are refreshed by plain copies where the multi-GPU run uses NVLink send/recv -- against the
whole-frame aggregation.  (The real multi-rank run is scripts/bandsplit_check.py under torchrun.)"""
import ctypes as C

import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.bandsplit import HALO, aggregate_cross_band, make_band
from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def test_single_band_equals_whole_frame(engine):
    import torch

    H, W, D = 90, 160, 40
    p = make_pair(H, W, D, 51, "scene")
    bgr = torch.as_tensor(p["bgr_left"], device="cuda:0")
    gl, gr = torch.as_tensor(p["gray_left"], device="cuda:0"), torch.as_tensor(p["gray_right"], device="cuda:0")
    cost = engine.ad_cost(gl, gr, D, 0)
    full = engine.aggregate_cross(bgr, cost, mode=SUM_FAST)["cost"]
    out = aggregate_cross_band(engine, bgr, cost.clone(), make_band(0, 1, H), iters=4, vmax=255.0)
    assert torch.equal(out, full)


@pytest.mark.parametrize("mode", [SUM_FAST, SUM_REFERENCE_ORDER], ids=["fast", "reference_order"])
@pytest.mark.parametrize("H,W,D,world", [(150, 120, 40, 2), (300, 96, 33, 3), (200, 70, 64, 4)])
def test_emulated_bands_match_whole_frame(engine, oracle, H, W, D, world, mode):
    import torch

    p = make_pair(H, W, D, 52 + world, "scene")
    dev = "cuda:0"
    bgr = torch.as_tensor(p["bgr_left"], device=dev)
    gl, gr = torch.as_tensor(p["gray_left"], device=dev), torch.as_tensor(p["gray_right"], device=dev)
    cost = engine.ad_cost(gl, gr, D, 0)
    full = engine.aggregate_cross(bgr, cost, mode=mode)["cost"]
    maps = engine.cross_arms(bgr)
    bands = [make_band(r, world, H) for r in range(world)]
    cur = [cost[b.b0:b.b1].clone() for b in bands]
    oth = [torch.empty_like(c) for c in cur]
    sl = [{k: maps[k][b.b0:b.b1].contiguous() for k in ("arms", "sup0", "sup1")} for b in bands]
    lib = engine.lib
    for s in range(lib.smt_cross_sweep_count(4)):
        horiz, div = C.c_int32(0), C.c_int32(0)
        lib.smt_cross_sweep_info(4, s, C.byref(horiz), C.byref(div))
        if not horiz.value:
            engine.sync()
            for r, b in enumerate(bands):   # what exchange_halos does with send/recv
                if r > 0:
                    up = bands[r - 1]
                    cur[r][0:b.top] = cur[r - 1][up.nrows - up.bottom - HALO:up.nrows - up.bottom]
                if r < world - 1:
                    dn = bands[r + 1]
                    cur[r][b.nrows - b.bottom:] = cur[r + 1][dn.top:dn.top + HALO]
            torch.cuda.synchronize()
        for r, b in enumerate(bands):
            engine._ck(lib.smt_aggregate_cross_sweep(engine.h, 4, s, C.c_void_p(cur[r].data_ptr()), C.c_void_p(oth[r].data_ptr()),
                                                     C.c_void_p(sl[r]["arms"].data_ptr()), C.c_void_p(sl[r]["sup0"].data_ptr()),
                                                     C.c_void_p(sl[r]["sup1"].data_ptr()), b.nrows, W, D, 255.0, int(mode), 1))
        cur, oth = oth, cur
    engine.sync()
    got = torch.cat([cur[r][b.top:b.top + b.r1 - b.r0] for r, b in enumerate(bands)], 0)
    # vertical prefixes start at different rows in a band, so the compensated sums may differ in the last bit
    same = float((got == full).float().mean())
    rel = ((got - full).abs() / full.abs().clamp_min(1e-30))
    rel[got == full] = 0
    if mode == SUM_REFERENCE_ORDER:
        assert torch.equal(got, full)        # band split == single GPU, bit for bit (test pyramid level 4)
    else:
        assert float(rel.max()) <= 2e-6, (same, float(rel.max()))
    # and the whole thing is the reference's CrossAggregator within the usual tolerance
    ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost.cpu().numpy())["cost"]
    g = got.cpu().numpy()
    r = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-30)
    r[g == ref] = 0
    assert r.max() <= (0.0 if mode == SUM_REFERENCE_ORDER else 1e-5)

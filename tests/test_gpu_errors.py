"""GPU: error behaviour of the C ABI -- every bad call returns a status and a message, nothing aborts or falls back."""
import ctypes as C

import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST, SmtError
from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def test_non_integer_image_is_rejected(engine):
    L = np.full((8, 16), 3.5, np.float32)
    with pytest.raises(SmtError) as e:
        engine.adcensus_cost(L, L, 4)
    assert e.value.status == 5 and "0..255" in str(e.value)
    with pytest.raises(SmtError):
        engine.adcensus_cost(np.full((8, 16), 300.0, np.float32), L * 0, 4)


def test_null_and_bad_sizes(engine):
    lib, h = engine.lib, engine.h
    assert lib.smt_wta_first_min(h, None, 4, 4, 4, None, 0) == 1
    assert lib.smt_scanline(h, None, None, 0, 0, 0, 10, 150, None, None, None, None, None, None, 0) == 1
    assert b"bad argument" in lib.smt_last_error_string(h)
    assert lib.smt_sync(None) == 1
    buf = np.zeros(16, np.float32)
    assert lib.smt_median(h, C.c_void_p(buf.ctypes.data), C.c_void_p(buf.ctypes.data), 4, 4, 4, 0) == 2   # even window


def test_unsupported_parameters(engine):
    p = make_pair(20, 40, 8, 1, "scene")
    cost = np.ones((20, 40, 8), np.float32)
    with pytest.raises(SmtError) as e:       # the fast path's halo assumes arms <= 34
        engine.aggregate_cross(p["bgr_left"], cost, L1=40, mode=SUM_FAST)
    assert e.value.status == 2
    engine.aggregate_cross(p["bgr_left"], cost, L1=40)   # the reference-order path has no such limit
    big = np.zeros((2, 2, 600), np.float32)
    with pytest.raises(SmtError) as e:       # one warp holds at most 512 disparities of a scanline
        engine.scanline(big, np.zeros((2, 2), np.float32))
    assert e.value.status == 2


def test_engine_survives_errors(engine, oracle):
    """After failed calls the context keeps working and stays correct."""
    p = make_pair(24, 48, 8, 2, "noise")
    L, R = p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32)
    with pytest.raises(SmtError):
        engine.adcensus_cost(L + 0.25, R, 8)
    got = engine.adcensus_cost(L, R, 8)["cost"]
    assert np.array_equal(got, oracle.adcensus(L, R, 8, want=("costL",))["costL"])

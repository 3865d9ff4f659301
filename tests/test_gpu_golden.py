"""GPU: the CUDA path against the COMMITTED golden fixtures (tests/golden/, generated from the compiled reference by
oracle/make_golden.py) -- no oracle library is loaded here, so this also pins parity on a box without oracle/_ref."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
DIGESTS = json.loads((GOLDEN / "digests.json").read_text())
CASES = [k for k in DIGESTS if k.startswith("adc_")]


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def engine_path(e, H, W, D, kind, seed):
    """Every stage of P_full through the C ABI (reference-order sums), each fed with the previous stage's output --
    the same chain oracle/make_golden.py:adcensus_path runs on the compiled reference."""
    p = make_pair(H, W, D, seed, kind)
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    out = {}
    for view, s in ((0, "L"), (1, "R")):
        c = e.adcensus_cost(L, R, D, 10.0, 30.0, view, want_ad=True, want_census=True)
        out["adcensus_cost" + s], out["adcensus_ad" + s], out["adcensus_census" + s] = c["cost"], c["ad"], c["census"]
        out["adcensus_wta" + s] = e.wta(c["cost"])
    for view, own in (("L", "bgr_left"), ("R", "bgr_right")):
        a = e.aggregate_cross(p[own], out["adcensus_cost" + view], want_arms=True)
        out[f"agg{view}_cost"], out[f"agg{view}_arms"] = a["cost"], a["arms"]
        out[f"agg{view}_sup0"], out[f"agg{view}_sup1"] = a["sup0"], a["sup1"]
    s = e.scanline(out["aggL_cost"], L, 10, 150, want=("left", "right", "up", "down", "sum", "disp"))
    out.update({f"scan_{k}": v for k, v in s.items()})
    out["dispR"] = e.wta(out["aggR_cost"])
    out["lrc_disp"], out["lrc_occ"], out["lrc_mis"] = e.lrcheck(s["disp"], out["dispR"], 2)
    out["lrand_last"], out["lrand_occ"], out["lrand_mis"] = e.lrcheck_and(s["disp"], out["dispR"], 2.0)
    out["speckle_inf"] = e.remove_speckles(out["lrc_disp"], 1, 30)
    out["speckle_m1"] = e.remove_speckles(s["disp"], 1, 12, -1)
    out["median3"] = e.median(out["speckle_inf"], 3)
    out["median5"] = e.median(s["disp"], 5)
    return out


@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_reference_digests(engine, case):
    meta = DIGESTS[case]
    H, W, D = meta["shape"]
    got = engine_path(engine, H, W, D, meta["kind"], meta["seed"])
    want = meta["outputs"]
    assert set(want) <= set(got), sorted(set(want) - set(got))
    bad = [k for k in want if digest(np.asarray(got[k])) != want[k]]
    assert not bad, f"{case}: {bad}"


def test_cuda_path_reproduces_golden_arrays(engine):
    for f in sorted(GOLDEN.glob("adc_*.npz")):
        z = np.load(f)
        meta = DIGESTS[f.stem]
        H, W, D = meta["shape"]
        got = engine_path(engine, H, W, D, meta["kind"], meta["seed"])
        for k in got:
            if k in z.files:
                assert np.array_equal(np.asarray(got[k]), z[k], equal_nan=True), f"{f.stem}: {k}"

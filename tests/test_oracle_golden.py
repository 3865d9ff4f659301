"""CPU: the oracle (C port, and the compiled reference when present) against the golden fixtures.

The fixtures are outputs of the compiled, unmodified reference (oracle/make_golden.py); this is what pins the port.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.make_golden import ADCENSUS_CASES, adcensus_path, digest

GOLDEN = Path(__file__).resolve().parent / "golden"
DIGESTS = json.loads((GOLDEN / "digests.json").read_text())


@pytest.mark.parametrize("case", ADCENSUS_CASES, ids=[c[0] for c in ADCENSUS_CASES])
def test_port_reproduces_reference_golden(oracle_port, case):
    name, H, W, D, kind, seed, full = case
    out = adcensus_path(oracle_port, H, W, D, kind, seed)
    want = DIGESTS[name]["outputs"]
    assert set(out) == set(want)
    for k, v in out.items():
        assert digest(v) == want[k], f"{name}:{k}"
    if full:
        z = np.load(GOLDEN / f"{name}.npz")
        for k, v in out.items():
            assert np.array_equal(v, z[k]), f"{name}:{k}"


@pytest.mark.skipif(not orc.have("ref"), reason="compiled reference (oracle/_ref) not present")
def test_compiled_reference_still_matches_golden():
    o = orc.Oracle("ref")
    name, H, W, D, kind, seed, _ = ADCENSUS_CASES[0]
    out = adcensus_path(o, H, W, D, kind, seed)
    for k, v in out.items():
        assert digest(v) == DIGESTS[name]["outputs"][k], k


def test_first_min_tie_break_and_constant_image(oracle_port):
    """All-ties input: first-minimum WTA must return 0 (AD-Census.h:359-372)."""
    vol = np.ones((4, 5, 9), np.float32)
    assert np.all(oracle_port.wta(vol) == 0)
    vol[..., 3] = 0.5
    vol[..., 6] = 0.5
    assert np.all(oracle_port.wta(vol) == 3)


# ---------------------------------------------------------------------------------------------- matchers / CrossArm / CBLSM
from oracle.make_golden_matchers import CROSSARM_CASES, MATCHER_CASES, crossarm_outputs, matcher_outputs  # noqa: E402


@pytest.mark.parametrize("case", MATCHER_CASES, ids=[c[0] for c in MATCHER_CASES])
def test_port_matchers_reproduce_reference_golden(oracle_port, case):
    name = case[0]
    out = matcher_outputs(oracle_port, *case[1:])
    for k, v in out.items():
        assert digest(v) == DIGESTS[name]["outputs"][k], f"{name}:{k}"
    f = GOLDEN / f"{name}.npz"
    if f.exists():
        z = np.load(f)
        for k, v in out.items():
            assert np.array_equal(v, z[k], equal_nan=v.dtype.kind == "f"), f"{name}:{k}"


@pytest.mark.parametrize("case", CROSSARM_CASES, ids=[c[0] for c in CROSSARM_CASES])
def test_port_crossarm_cblsm_reproduce_reference_golden(oracle_port, case):
    name = case[0]
    out = crossarm_outputs(oracle_port, *case[1:])
    for k, v in out.items():
        assert digest(v) == DIGESTS[name]["outputs"][k], f"{name}:{k}"


def test_sad_optimal_disparity_quirks(oracle_port):
    """Sad.h:40-85: the search starts at d=1, d=0 wins ties, the last disparity is rejected."""
    assert oracle_port.sad_optimal_disparity(np.array([5, 9, 3, 7], np.float32)) == 2
    assert oracle_port.sad_optimal_disparity(np.array([3, 9, 3, 7], np.float32)) == 0      # sad[0] <= min
    assert oracle_port.sad_optimal_disparity(np.array([5, 9, 7, 3], np.float32)) == 0      # best == D-1
    assert oracle_port.sad_optimal_disparity(np.array([9, 4, 4, 7], np.float32)) == 1      # first minimum


def test_port_transform_to_show_matches_reference_golden(oracle_port):
    z = np.load(GOLDEN / "prepost.npz")
    assert np.array_equal(oracle_port.transform_to_show(z["in_disp"]), z["show"])


def test_port_fill_the_hole_matches_reference_golden(oracle_port):
    """FillTheHole as written (PostProcessing.h:156): fixtures from the compiled reference (oracle/make_golden_matchers.py)."""
    from oracle.make_golden_matchers import FILL_CASES, fill_inputs

    z = np.load(GOLDEN / "fill.npz")
    for H, W, D, seed in FILL_CASES:
        d, occ, mis = fill_inputs(H, W, D, seed)
        f, m = oracle_port.fill_the_hole(d, D, occ, mis)
        assert np.array_equal(f, z[f"fill_{H}x{W}x{D}"]) and np.array_equal(m, z[f"mis_{H}x{W}x{D}"])
    # phase 3 is skipped when the mismatch list is empty (:175), and nothing happens with both lists empty
    d, occ, _ = fill_inputs(20, 28, 8, 3)
    f, m = oracle_port.fill_the_hole(d, 8, occ, np.zeros((0, 2), np.int32))
    assert len(m) == 0 and (f == 65535).sum() > 0
    f, _ = oracle_port.fill_the_hole(d, 8, np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32))
    assert np.array_equal(f, d)


def test_port_display_chain_matches_cv2_and_reference_golden(oracle_port):
    """ASW display chain (ASWeight.cpp:66-78): the three OpenCV steps against cv2 4.13 outputs, FillImageNew against the
    compiled reference (oracle/make_golden_display.py)."""
    z = np.load(GOLDEN / "display.npz")
    n = sum(1 for f in z.files if f.startswith("in"))
    assert n >= 7
    for k in range(n):
        a = z[f"in{k}"]
        assert np.array_equal(oracle_port.normalize_minmax_u8_from_u8(a), z[f"norm{k}"]), k
        for tag, (nv, ms, md) in {"a": (0, 40, 2), "b": (0, 5, 0), "c": (7, 100, 10)}.items():
            assert np.array_equal(oracle_port.filter_speckles_u8(a, nv, ms, md), z[f"speck{tag}{k}"]), (k, tag)
        assert np.array_equal(oracle_port.median_blur_u8(a, 3), z[f"med3_{k}"]), k
        assert np.array_equal(oracle_port.median_blur_u8(a, 5), z[f"med5_{k}"]), k
        for pe in (0, 9):
            if f"fill{pe}_{k}" in z.files:
                assert np.array_equal(oracle_port.fill_image_new(a, pe)[0], z[f"fill{pe}_{k}"]), (k, pe)

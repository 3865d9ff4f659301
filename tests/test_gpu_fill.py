"""GPU: smt_fill_the_hole against FillTheHole as written (AD-CensusV1/PostProcessing.h:156-248): golden values from the
compiled reference, the oracle port on larger inputs (also where the reference itself overflows its fill vector)."""
from pathlib import Path

import numpy as np
import pytest

from oracle.make_golden_matchers import FILL_CASES, fill_inputs

pytestmark = pytest.mark.gpu
Z = np.load(Path(__file__).resolve().parent / "golden" / "fill.npz")


@pytest.mark.parametrize("case", FILL_CASES, ids=lambda c: "x".join(map(str, c[:3])))
def test_fill_matches_reference_golden(engine, case):
    H, W, D, seed = case
    d, occ, mis = fill_inputs(H, W, D, seed)
    f, m = engine.fill_the_hole(d, D, occ, mis)
    assert np.array_equal(f, Z[f"fill_{H}x{W}x{D}"])
    assert np.array_equal(m, Z[f"mis_{H}x{W}x{D}"])


@pytest.mark.parametrize("shape", [(96, 130, 32), (130, 96, 24), (64, 64, 64)])
def test_fill_matches_port_many_holes(engine, oracle_port, shape):
    """More 0xffff pixels than mismatches: the reference overflows (UB), the port and the kernel fill all of them."""
    H, W, D = shape
    d, occ, mis = fill_inputs(H, W, D, 11, frac_ffff=0.3, npx=400)
    f, m = engine.fill_the_hole(d, D, occ, mis)
    fo, mo = oracle_port.fill_the_hole(d, D, occ, mis)
    assert np.array_equal(f, fo) and np.array_equal(m, mo)
    assert len(m) > len(mis)


def test_fill_unsorted_lists_with_duplicates(engine, oracle_port):
    """List order matters twice: the direction table switches at the FIRST entry with y == height/2, and the last of
    several entries for one pixel wins (:198-200, :239-244)."""
    H, W, D = 40, 52, 16
    d, occ, mis = fill_inputs(H, W, D, 21, frac_ffff=0.05, npx=300)
    rng = np.random.default_rng(4)
    occ = np.concatenate([occ, occ[:50]])[rng.permutation(len(occ) + 50)]
    mis = np.concatenate([mis, mis[-50:]])[rng.permutation(len(mis) + 50)]
    f, m = engine.fill_the_hole(d, D, occ, mis)
    fo, mo = oracle_port.fill_the_hole(d, D, occ, mis)
    assert np.array_equal(f, fo) and np.array_equal(m, mo)


def test_fill_empty_lists_and_device_arrays(engine, oracle_port):
    import torch

    H, W, D = 30, 44, 12
    d, occ, mis = fill_inputs(H, W, D, 3)
    none = np.zeros((0, 2), np.int32)
    f, m = engine.fill_the_hole(d, D, occ, none)                   # phase 3 skipped: old mismatch list empty (:175)
    fo, mo = oracle_port.fill_the_hole(d, D, occ, none)
    assert np.array_equal(f, fo) and len(m) == 0 and len(mo) == 0
    f, _ = engine.fill_the_hole(d, D, none, none)
    assert np.array_equal(f, d)
    ft, mt = engine.fill_the_hole(torch.as_tensor(d, device="cuda"), D, occ, mis)
    fo, mo = oracle_port.fill_the_hole(d, D, occ, mis)
    assert np.array_equal(ft.cpu().numpy(), fo) and np.array_equal(mt.cpu().numpy(), mo)


def test_fill_refuses_entries_outside_its_frame(engine):
    from stereo_match_traditional_b200.api import SmtError

    d = np.zeros((10, 20), np.float32)
    bad = np.array([[5, 15]], np.int32)      # x = 15 >= width = rows = 10
    with pytest.raises(SmtError):
        engine.fill_the_hole(d, 8, bad, np.zeros((0, 2), np.int32))


def test_fill_as_written_refuses_tall_frame_lists_untouched(engine):
    """rows > cols: an ordinary LeftRightConsistency list holds y >= cols, which the as-written variant (y < col, x < row,
    PostProcessing.h:158-159) cannot address.  Also negative entries.  The call must fail with INVALID_ARG before the map
    is read or written through such an entry, and the context must stay usable."""
    from stereo_match_traditional_b200.api import SmtError

    rows, cols = 40, 12
    rng = np.random.default_rng(5)
    d = rng.integers(0, 8, (rows, cols)).astype(np.float32)
    occ = np.array([[3, 4], [30, 5]], np.int32)           # y = 30 >= height = cols = 12
    mis = np.array([[2, 2], [-7, 1]], np.int32)           # negative row
    for o, m in ((occ, np.zeros((0, 2), np.int32)), (occ[:1], mis)):
        before = d.copy()
        with pytest.raises(SmtError):
            engine.fill_the_hole(d, 8, o, m)
        assert np.array_equal(d, before)
    ok, _ = engine.fill_the_hole(d, 8, occ[:1], np.zeros((0, 2), np.int32))   # the context still works
    assert ok.shape == d.shape


@pytest.mark.parametrize("shape", [(48, 80, 16), (80, 48, 24), (120, 200, 32)])
def test_fill_fixed_variant_after_lrcheck(engine, oracle_port, shape):
    """SMT_FILL_FIXED (axes not swapped, invalid = +inf) fed by the lists and the +inf map of LeftRightConsistency.
    There is no reference implementation of this variant: the checker is the port with the same two changes."""
    H, W, D = shape
    rng = np.random.default_rng(17)
    truth = np.clip((np.add.outer(np.arange(H), np.arange(W)) // 9) % D, 0, D - 1).astype(np.float32)
    dl = truth.copy()
    dr = truth.copy()
    bad = rng.random((H, W)) < 0.15
    dl[bad] = rng.integers(0, D, bad.sum()).astype(np.float32)
    chk, occ, mis = engine.lrcheck(dl, dr, 2)
    assert np.isinf(chk).sum() == len(occ) + len(mis) > 0
    f, m = engine.fill_the_hole(chk, D, occ, mis, fixed=True)
    fo, mo = oracle_port.fill_the_hole(chk, D, occ, mis, fixed=True)
    assert np.array_equal(f, fo) and np.array_equal(m, mo)
    assert np.isinf(f).sum() == 0 and len(m) == 0          # every hole got a value; nothing left for phase 3

"""GPU: the ASW driver's display chain (ASW/ASWeight.cpp:66-78) through the C ABI.

normalize / filterSpeckles / medianBlur: bit-exact against cv2 4.13 outputs (tests/golden/display.npz) and against the C
port at larger sizes; FillImageNew: bit-exact against the compiled reference's outputs in the same fixture, including the
index shift the reference produces when a search ends on column == cols."""
from pathlib import Path

import numpy as np
import pytest

from oracle.make_golden_display import disparity_like

pytestmark = pytest.mark.gpu
Z = np.load(Path(__file__).resolve().parent / "golden" / "display.npz")
NCASE = sum(1 for f in Z.files if f.startswith("in"))
SPECK = {"a": (0, 40, 2), "b": (0, 5, 0), "c": (7, 100, 10)}


@pytest.mark.parametrize("k", range(NCASE))
def test_display_steps_against_golden(engine, k):
    a = Z[f"in{k}"]
    assert np.array_equal(engine.normalize_minmax_u8_from_u8(a), Z[f"norm{k}"])
    for tag, (nv, ms, md) in SPECK.items():
        assert np.array_equal(engine.filter_speckles_u8(a, nv, ms, md), Z[f"speck{tag}{k}"]), tag
    assert np.array_equal(engine.median_blur_u8(a, 3), Z[f"med3_{k}"])
    assert np.array_equal(engine.median_blur_u8(a, 5), Z[f"med5_{k}"])
    for pe in (0, 9):
        if f"fill{pe}_{k}" in Z.files:
            got, nun = engine.fill_image_new(a, pe)
            assert np.array_equal(got, Z[f"fill{pe}_{k}"]), pe


@pytest.mark.parametrize("shape,seed", [((375, 450), 11), ((720, 1280), 12), ((1080, 1920), 13), ((33, 2000), 14), ((3, 3), 15)])
def test_display_steps_against_port(engine, oracle_port, shape, seed):
    a = disparity_like(shape[0], shape[1], seed, holes=40)
    assert np.array_equal(engine.normalize_minmax_u8_from_u8(a), oracle_port.normalize_minmax_u8_from_u8(a))
    for nv, ms, md in SPECK.values():
        assert np.array_equal(engine.filter_speckles_u8(a, nv, ms, md), oracle_port.filter_speckles_u8(a, nv, ms, md))
    for ks in (3, 5):
        assert np.array_equal(engine.median_blur_u8(a, ks), oracle_port.median_blur_u8(a, ks))
    want, nun_want = oracle_port.fill_image_new(a, 0)
    got, nun = engine.fill_image_new(a, 0)
    assert nun == nun_want and np.array_equal(got, want)


def test_fill_image_new_value_shift(engine, oracle_port):
    """A search that ends exactly on column == cols with a zero there records no value (ASW.h:485): later pixels take
    their successor's value.  Row 0 col 2 probes 2, 3 == cols -> (1, 0) == 0 -> nothing recorded."""
    a = np.array([[0, 0, 0], [0, 5, 0], [7, 0, 0]], np.uint8)
    want, nun_want = oracle_port.fill_image_new(a, 0)
    got, nun = engine.fill_image_new(a, 0)
    assert nun == nun_want and nun > 0
    assert np.array_equal(got, want)
    got9, _ = engine.fill_image_new(a, 9)       # the byte below the last row is part of the input
    assert np.array_equal(got9, oracle_port.fill_image_new(a, 9)[0])


def test_constant_and_all_zero_images(engine, oracle_port):
    z = np.zeros((9, 14), np.uint8)
    c = np.full((9, 14), 3, np.uint8)
    assert np.all(engine.normalize_minmax_u8_from_u8(c) == 0)              # scale 0
    assert np.array_equal(engine.filter_speckles_u8(c, 0, 40, 2), c)        # one component of 126 pixels > 40
    assert np.all(engine.filter_speckles_u8(c, 0, 126, 2) == 0)             # <= maxSpeckleSize goes
    got, nun = engine.fill_image_new(z, 0)
    want, nun_want = oracle_port.fill_image_new(z, 0)
    assert np.array_equal(got, want) and nun == nun_want
    assert np.array_equal(engine.median_blur_u8(z, 5), z)


@pytest.mark.parametrize("k", [0, 1, 3])
def test_display_chain_equals_its_steps(engine, k):
    """smt_asw_display_chain = CrossCheckDiaparity + the four steps with the driver's constants; outputs are the three
    images the driver writes."""
    a = Z[f"in{k}"]
    dl = a.astype(np.float32)
    dr = np.roll(dl, -3, axis=1) + (np.random.default_rng(k).random(a.shape) < 0.1) * 9.0
    dr = dr.astype(np.float32)
    from_cc = engine.asw_crosscheck(dl, dr)
    spk, fil, med = engine.asw_display_chain(dl, dr)
    s = engine.median_blur_u8(engine.filter_speckles_u8(engine.normalize_minmax_u8_from_u8(from_cc), 0, 40, 2), 5)
    f, _ = engine.fill_image_new(s, 0)
    assert np.array_equal(spk, s) and np.array_equal(fil, f) and np.array_equal(med, engine.median_blur_u8(f, 3))


def test_display_chain_device_memspace(engine, oracle_port):
    import torch

    a = disparity_like(120, 300, 21, holes=12)
    dl = a.astype(np.float32)
    dr = dl.copy()
    t_l, t_r = torch.as_tensor(dl, device="cuda"), torch.as_tensor(dr, device="cuda")
    spk_d, fil_d, med_d = engine.asw_display_chain(t_l, t_r)
    spk_h, fil_h, med_h = engine.asw_display_chain(dl, dr)
    assert np.array_equal(spk_d.cpu().numpy(), spk_h) and np.array_equal(fil_d.cpu().numpy(), fil_h) and np.array_equal(med_d.cpu().numpy(), med_h)
    # against the oracle, step by step (the cross-check itself is covered in test_gpu_matchers.py; here j - d may leave the
    # row, where the reference reads the neighbouring row)
    cc = engine.asw_crosscheck(dl, dr)
    s = oracle_port.median_blur_u8(oracle_port.filter_speckles_u8(oracle_port.normalize_minmax_u8_from_u8(cc), 0, 40, 2), 5)
    f, _ = oracle_port.fill_image_new(s, 0)
    assert np.array_equal(spk_h, s) and np.array_equal(fil_h, f) and np.array_equal(med_h, oracle_port.median_blur_u8(f, 3))


def test_display_argument_errors(engine):
    from stereo_match_traditional_b200.api import SmtError

    a = np.ones((4, 4), np.uint8)
    with pytest.raises(SmtError):
        engine.median_blur_u8(a, 7)

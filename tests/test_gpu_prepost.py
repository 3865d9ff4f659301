"""GPU: the drivers' pre/post steps against golden values: TransformToShow from the compiled reference,
cvtColor / copyMakeBorder / normalize from the cv2 wheel of the build container (tests/golden/prepost.npz)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Z = np.load(Path(__file__).resolve().parent / "golden" / "prepost.npz")


def test_bgr_to_gray_and_padding(engine):
    assert np.array_equal(engine.bgr_to_gray(Z["in_bgr"]), Z["gray"])
    assert np.array_equal(engine.replicate_pad(Z["gray"], 4), Z["pad4"])
    assert np.array_equal(engine.u8_to_f32(Z["gray"]), Z["gray"].astype(np.float32))


def test_transform_to_show(engine, oracle):
    assert np.array_equal(engine.transform_to_show(Z["in_disp"]), Z["show"])
    assert np.array_equal(engine.transform_to_show(Z["in_disp"]), oracle.transform_to_show(Z["in_disp"]))


def test_normalize_minmax(engine):
    # main.cpp:106-107 normalises the 8-bit display image; SADmain.cpp:87 a wider one
    got = engine.normalize_minmax_u8(Z["show"].astype(np.float32))
    assert np.array_equal(got, Z["norm_show"])
    got = engine.normalize_minmax_u8(Z["in_smooth"]).astype(np.int32)
    ref = Z["norm_smooth"].astype(np.int32)
    # OpenCV rounds src*scale+shift in float; one grey level of difference is tolerated at exact .5 ties
    assert np.abs(got - ref).max() <= 1 and (got == ref).mean() >= 0.99
    assert np.all(engine.normalize_minmax_u8(np.full((4, 5), 7.0, np.float32)) == 0)   # constant image: scale 0


def test_driver_flow_raw_bgr_to_sad(engine, oracle):
    """SADmain.cpp:40-66 with everything on the device: BGR -> gray -> replicate pad -> GetPointDepthLeft."""
    from stereo_match_traditional_b200.synth import make_pair, replicate_pad

    p = make_pair(40, 90, 16, 9, "scene")
    gl, gr = engine.bgr_to_gray(p["bgr_left"]), engine.bgr_to_gray(p["bgr_right"])
    assert np.array_equal(gl, p["gray_left"]) and np.array_equal(gr, p["gray_right"])
    lp, rp = engine.replicate_pad(gl, 4), engine.replicate_pad(gr, 4)
    assert np.array_equal(engine.sad(lp, rp, 16, 3), oracle.sad(replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4), 16, 3))


def _read_png(path):
    """Minimal PNG reader (8-bit grey / RGB, filter type 0 rows only): no dependency on an image library."""
    import struct
    import zlib

    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    at, idat, hdr = 8, b"", None
    while at < len(b):
        n, typ = struct.unpack(">I4s", b[at:at + 8])
        data = b[at + 8:at + 8 + n]
        assert struct.unpack(">I", b[at + 8 + n:at + 12 + n])[0] == zlib.crc32(typ + data) & 0xffffffff
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        at += 12 + n
    W, H, depth, ctype = hdr[:4]
    ch = {0: 1, 2: 3}[ctype]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, W * ch + 1)
    assert depth == 8 and np.all(raw[:, 0] == 0)
    return raw[:, 1:].reshape(H, W, ch) if ch == 3 else raw[:, 1:]


def test_write_png(engine, tmp_path):
    """imwrite of the drivers' 8-bit results (main.cpp:112-114): pixel-exact, readable by any PNG reader."""
    rng = np.random.default_rng(5)
    gray = rng.integers(0, 256, (375, 450), dtype=np.uint8)       # > 65535 bytes: several stored deflate blocks
    bgr = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    engine.write_png(tmp_path / "g.png", gray)
    engine.write_png(tmp_path / "c.png", bgr)
    assert np.array_equal(_read_png(tmp_path / "g.png"), gray)
    assert np.array_equal(_read_png(tmp_path / "c.png"), bgr[:, :, ::-1])   # stored as RGB
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(cv2.imread(str(tmp_path / "g.png"), cv2.IMREAD_UNCHANGED), gray)
    assert np.array_equal(cv2.imread(str(tmp_path / "c.png"), cv2.IMREAD_UNCHANGED), bgr)
    import torch

    engine.write_png(tmp_path / "d.png", torch.as_tensor(gray, device="cuda"))   # device memspace
    assert np.array_equal(_read_png(tmp_path / "d.png"), gray)

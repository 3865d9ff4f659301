"""ORACLE tooling: golden fixtures for the ASW driver's display chain (ASW/ASWeight.cpp:66-78) -> tests/golden/display.npz.

normalize / filterSpeckles / medianBlur are OpenCV functions; their golden outputs come from the cv2 wheel of the build
container (4.13), which does not exist on the GPU box.  FillImageNew comes from the compiled reference (oracle/_ref).
Run from the repo root:  python oracle/make_golden_display.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def disparity_like(H, W, seed, holes=6):
    """An 8-bit map that looks like a cross-checked disparity image: a smooth ramp with noise, isolated speckles, zero
    blocks (failed cross-checks), a zero band at the left border and sometimes whole zero rows."""
    r = np.random.default_rng(seed)
    base = np.linspace(5, 60, W)[None, :] + 10 * np.sin(np.arange(H) / 7.0)[:, None] + r.normal(0, 1.2, (H, W))
    img = np.clip(base, 0, 255).astype(np.uint8)
    m = r.random((H, W)) < 0.05
    img[m] = r.integers(0, 255, int(m.sum()))
    for _ in range(holes):
        y, x = int(r.integers(0, H)), int(r.integers(0, W))
        img[y:y + int(r.integers(1, 12)), x:x + int(r.integers(1, 30))] = 0
    img[:, : int(r.integers(0, 12))] = 0
    if seed % 3 == 0:
        img[int(r.integers(0, H))] = 0
    return img


CASES = [(37, 53, 1), (64, 200, 2), (5, 7, 3), (120, 90, 4), (1, 40, 5), (40, 1, 6), (80, 333, 9)]


def main():
    import cv2

    from oracle import oracle as orc

    ref, port = orc.Oracle("ref"), orc.Oracle("port")

    def ref_fill(img, past_end):
        """The reference indexes its value vector past the end for the last `n_unassigned` zero pixels (a crash when the
        vector is empty or the overrun is long): only run it where the port predicts a short overrun, and blank those
        pixels in the fixture (they hold whatever the heap held)."""
        mine, nun = port.fill_image_new(img, past_end)
        if nun > 16 or nun >= int((img == 0).sum()):
            return None
        got = ref.fill_image_new(img, past_end)[0]
        if nun:
            z = np.flatnonzero(img.ravel() == 0)[-nun:]
            got.ravel()[z] = 0
        assert np.array_equal(got, mine), "port and compiled reference disagree on FillImageNew"
        return got
    out = {"cv2_version": np.array(cv2.__version__)}
    for k, (H, W, seed) in enumerate(CASES):
        a = disparity_like(H, W, seed)
        out[f"in{k}"] = a
        out[f"norm{k}"] = cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX)
        for tag, (nv, ms, md) in {"a": (0, 40, 2), "b": (0, 5, 0), "c": (7, 100, 10)}.items():
            b = a.copy()
            cv2.filterSpeckles(b, nv, ms, md)
            out[f"speck{tag}{k}"] = b
        if min(H, W) >= 1:
            out[f"med3_{k}"] = cv2.medianBlur(a, 3)
            out[f"med5_{k}"] = cv2.medianBlur(a, 5)
        for pe in (0, 9):
            f = ref_fill(a, pe)
            if f is not None:
                out[f"fill{pe}_{k}"] = f
        # the chain on the 8-bit map, OpenCV for the three OpenCV steps and the compiled reference for FillImageNew
        c = cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX)
        cv2.filterSpeckles(c, 0, 40, 2)
        c = cv2.medianBlur(c, 5)
        out[f"chain_speckle{k}"] = c
        c = ref_fill(c, 0)
        if c is None:
            continue
        out[f"chain_filled{k}"] = c
        out[f"chain_median{k}"] = cv2.medianBlur(c, 3)
    np.savez_compressed(ROOT / "tests" / "golden" / "display.npz", **out)
    print("wrote tests/golden/display.npz", len(out), "arrays")


if __name__ == "__main__":
    main()

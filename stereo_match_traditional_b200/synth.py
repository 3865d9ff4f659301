"""Synthetic rectified stereo pairs.

The reference ships no images (its drivers load im2.png / im6.png that are not in the
tree, AD-CensusV1/main.cpp:16-17), so every parity test and the benchmark use seeded
synthetic 8-bit pairs.  Three value distributions:

  * ``scene``    piecewise-smooth scene with planted integer disparities (arms reach the
                 34-pixel cap, so the sticky-threshold path of CrossArm.cpp:168-173 triggers)
  * ``noise``    i.i.d. uniform 0..255 (arms ~0-1, tie stress)
  * ``constant`` L = R = 128 (all ties: first-minimum WTA must give 0; NCC zero-variance path)
"""
from __future__ import annotations

import numpy as np

__all__ = ["bgr_to_gray", "make_pair", "replicate_pad"]


def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    """OpenCV's 8-bit BGR2GRAY fixed-point formula (what cvtColor does in every reference main)."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8)


def replicate_pad(img: np.ndarray, pad: int) -> np.ndarray:
    """copyMakeBorder(..., BORDER_REPLICATE) as in SADmain.cpp:47-48 / ASWeight.cpp:54-55."""
    return np.ascontiguousarray(np.pad(img, ((pad, pad), (pad, pad)) + ((0, 0),) * (img.ndim - 2), mode="edge"))


def make_pair(height: int, width: int, max_disp: int, seed: int, kind: str = "scene", frame: int = 0):
    """Return a dict with bgr_left, bgr_right (H,W,3 u8), gray_left, gray_right (H,W u8) and disp_gt (H,W i32)."""
    rng = np.random.default_rng(seed)
    H, W, D = int(height), int(width), int(max_disp)
    if kind == "constant":
        bl = np.full((H, W, 3), 128, np.uint8)
        br = bl.copy()
        gt = np.zeros((H, W), np.int32)
    elif kind == "noise":
        bl = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        br = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        gt = np.zeros((H, W), np.int32)
    elif kind == "scene":
        yy, xx = np.mgrid[0:H, 0:W]
        base = 60.0 + 0.15 * xx * (450.0 / W) + 0.3 * yy * (375.0 / H)
        band = (xx >= int(0.37 * W)) & (xx < int(0.69 * W))
        shift = frame % 16
        rx0, rx1 = int(0.12 * W) + shift, int(0.87 * W) + shift
        rect = (yy >= int(0.42 * H)) & (yy < int(0.75 * H)) & (xx >= rx0) & (xx < rx1)
        val = base + 70.0 * band - 45.0 * rect + rng.integers(-2, 3, (H, W))
        ph, pw = max(1, int(0.2 * H)), max(1, int(0.2 * W))
        py, px = int(0.1 * H), int(0.3 * W)
        val[py:py + ph, px:px + pw] = rng.integers(0, 256, (ph, pw))
        b = np.clip(np.rint(val), 0, 255)
        g = np.clip(b + rng.integers(-3, 4, (H, W)), 0, 255)
        r = np.clip(b + rng.integers(-3, 4, (H, W)), 0, 255)
        bl = np.stack([b, g, r], axis=-1).astype(np.uint8)
        d_bg = min(max(int(round(0.08 * D)), 0), max(D - 1, 0))
        d_fg = min(max(int(round(0.17 * D)), 0), max(D - 1, 0))
        gt = np.where(rect, d_fg, d_bg).astype(np.int32)
        src = np.minimum(W - 1, xx + gt)
        br = bl[yy, src].astype(np.int64) + rng.integers(-1, 2, (H, W, 1))
        br = np.clip(br, 0, 255).astype(np.uint8)
    else:
        raise ValueError(f"unknown kind {kind!r}")
    bl = np.ascontiguousarray(bl)
    br = np.ascontiguousarray(br)
    return {
        "bgr_left": bl,
        "bgr_right": br,
        "gray_left": bgr_to_gray(bl),
        "gray_right": bgr_to_gray(br),
        "disp_gt": gt,
    }

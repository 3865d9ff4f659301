"""ORACLE tooling: golden fixtures for the CrossArm / CBLSM / SAD / NCC / ASW paths (called by oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np

from stereo_match_traditional_b200.synth import make_pair, replicate_pad

# (name, H, W, D, kind, seed)
MATCHER_CASES = [
    ("mat_scene_30x80x16", 30, 80, 16, "scene", 201),
    ("mat_noise_24x60x40", 24, 60, 40, "noise", 202),
    ("mat_const_20x50x8", 20, 50, 8, "constant", 203),
]
CROSSARM_CASES = [
    ("ca_scene_48x200x4", 48, 200, 4, "scene", 211),
    ("ca_noise_50x220x3", 50, 220, 3, "noise", 212),
]


def matcher_outputs(o, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    gl, gr = p["gray_left"], p["gray_right"]
    out = {}
    lp, rp = replicate_pad(gl, 4), replicate_pad(gr, 4)     # SAD winsize 3 -> pad 4 (SADmain.cpp:33-48)
    out["sad_left"] = o.sad(lp, rp, D, 3)
    out["sad_right"] = o.sad(lp, rp, D, 3, True)
    out["sad_costs_left"] = o.sad_costs_left(lp, rp, D, 3)
    out["ncc_w4"] = o.ncc(gl, gr, D, 4)
    out["ncc_costs_w4"] = o.ncc_costs(gl, gr, D, 4)         # ComputeCost per (pixel, d), NaN on zero-variance windows
    la, ra = replicate_pad(gl, 4), replicate_pad(gr, 4)     # ASW winSize 3 -> pad 4
    out["asw_left"] = o.asw(la, ra, D, 3, 50, 30, 40)
    out["asw_costs_left"] = o.asw_costs(la, ra, D, 3, 50, 30, 40)
    asw_r = o.asw(la, ra, D, 3, 50, 30, 40, True)
    out["asw_right_defined"] = np.ascontiguousarray(asw_r[:, : W - 3 - 2])   # ASW.h:401 reads costVolume[-1] beyond
    out["asw_crosscheck"] = o.asw_crosscheck(out["asw_left"], np.zeros_like(out["asw_left"]))
    return out


def crossarm_outputs(o, H, W, D, kind, seed):
    p = make_pair(H, W, D, seed, kind)
    out = {}
    vol = np.random.default_rng(seed).random((H, W, D), dtype=np.float32)
    for tag, img in (("gray", p["gray_left"]), ("bgr", p["bgr_left"])):
        a = o.crossarm(img, D, 30, 1, vol)
        for k in ("left", "right", "top", "bottom"):
            out[f"crossarm_{tag}_{k}"] = a[k]
        out[f"crossarm_{tag}_tao"] = np.array([a["tao_final"]], np.int32)
        out[f"crossarm_{tag}_aggv"] = a["vol"]
        c = o.cblsm_arms(img, 25, 34, 17)
        for k in ("left", "right", "up", "down"):
            out[f"cblsm_{tag}_{k}"] = c[k]
    adL, adR = o.cblsm_ad(p["gray_left"], p["gray_right"], D)
    out["cblsm_adL"], out["cblsm_adR"] = adL, adR
    arms = o.cblsm_arms(p["gray_left"], 25, 34, 17)
    v1 = o.cblsm_aggregate_v5(adL, arms)
    out["cblsm_v5_1"] = v1
    out["cblsm_v5_2"] = o.cblsm_aggregate_v5(v1, arms)
    out["cblsm_wta"] = o.cblsm_wta(out["cblsm_v5_2"])
    return out


def prepost_outputs(o):
    """Pre/post steps of the drivers.  cvtColor / copyMakeBorder / normalize are OpenCV (third-party, 3.1.0 pinned by the
    .vcxproj files, not in the tree): their golden values come from the cv2 4.13 wheel in the build container."""
    import cv2

    rng = np.random.default_rng(77)
    bgr = rng.integers(0, 256, (23, 37, 3), dtype=np.uint8)
    # OpenCV 3.1.0 (the reference's version): RGB2Gray<uchar> with yuv_shift = 14, B2Y/G2Y/R2Y = 1868/9617/4899 and
    # CV_DESCALE rounding.  cv2 >= 4 switched to a 15-bit variant (3735/19235/9798) that differs by one grey level on
    # ~0.4 % of random pixels, so the 3.1.0 formula is restated here instead of calling the 4.13 wheel.
    b, g, r = (bgr[..., k].astype(np.int64) for k in range(3))
    gray = ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8)
    assert (gray == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).mean() > 0.99
    disp = rng.integers(0, 60, (23, 37)).astype(np.float32)
    disp[rng.random((23, 37)) < 0.1] = np.inf
    disp[rng.random((23, 37)) < 0.05] = 2147483648.0
    disp[0, :5] = [0, 255, 256, 300.7, -3.2]
    show = o.transform_to_show(disp)
    smooth = (rng.random((23, 37)) * 50 + 3).astype(np.float32)
    return {
        "in_bgr": bgr, "gray": gray, "pad4": cv2.copyMakeBorder(gray, 4, 4, 4, 4, cv2.BORDER_REPLICATE),
        "in_disp": disp, "show": show,
        "norm_show": cv2.normalize(show, None, 0, 255, cv2.NORM_MINMAX),
        "in_smooth": smooth,
        "norm_smooth": cv2.normalize(smooth, None, 0, 255, cv2.NORM_MINMAX).round().clip(0, 255).astype(np.uint8),
        "norm_smooth_cvt": cv2.convertScaleAbs(cv2.normalize(smooth, None, 0, 255, cv2.NORM_MINMAX)),
    }


FILL_CASES = [(30, 30, 16, 5), (24, 40, 12, 6), (40, 24, 20, 7), (33, 57, 8, 8)]


def fill_inputs(H, W, D, seed, frac_ffff=0.03, npx=120):
    """Inputs of FillTheHole in its own (axis-swapped) frame: y < W, x < H; few 0xffff pixels so that the reference's
    third phase stays inside its fill vector (PostProcessing.h:177 vs :188)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, D, (H, W)).astype(np.float32)
    d[rng.random((H, W)) < frac_ffff] = 65535.0
    d[rng.random((H, W)) < 0.05] = np.inf
    lists = []
    for _ in range(2):
        l = np.stack([rng.integers(0, W, npx), rng.integers(0, H, npx)], 1).astype(np.int32)
        lists.append(l[np.lexsort((l[:, 1], l[:, 0]))])
    return d, lists[0], lists[1]


def fill_outputs(o):
    out = {}
    for H, W, D, seed in FILL_CASES:
        d, occ, mis = fill_inputs(H, W, D, seed)
        assert (d == 65535).sum() <= len(mis)
        f, m = o.fill_the_hole(d, D, occ, mis)
        out[f"fill_{H}x{W}x{D}"] = f
        out[f"mis_{H}x{W}x{D}"] = m
    return out


def generate(o, golden_dir, digests, digest):
    np.savez_compressed(golden_dir / "fill.npz", **fill_outputs(o))
    np.savez_compressed(golden_dir / "prepost.npz", **prepost_outputs(o))
    for name, H, W, D, kind, seed in MATCHER_CASES:
        out = matcher_outputs(o, H, W, D, kind, seed)
        digests[name] = {"shape": [H, W, D], "kind": kind, "seed": seed, "outputs": {k: digest(v) for k, v in out.items()}}
    for name, H, W, D, kind, seed in CROSSARM_CASES:
        out = crossarm_outputs(o, H, W, D, kind, seed)
        digests[name] = {"shape": [H, W, D], "kind": kind, "seed": seed, "outputs": {k: digest(v) for k, v in out.items()}}
    small = matcher_outputs(o, *MATCHER_CASES[0][1:])
    np.savez_compressed(golden_dir / f"{MATCHER_CASES[0][0]}.npz", **small)

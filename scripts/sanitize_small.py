"""Small end-to-end exercise of every C-ABI entry point (for compute-sanitizer memcheck runs)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from stereo_match_traditional_b200.api import Engine, SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.synth import make_pair, replicate_pad

e = Engine(0)
for (H, W, D) in [(40, 100, 24), (37, 71, 33)]:
    p = make_pair(H, W, D, 5, "scene")
    L, R = p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32)
    for mode in (SUM_REFERENCE_ORDER, SUM_FAST):
        e.pipeline(p["bgr_left"], p["bgr_right"], D, p["gray_left"], p["gray_right"], sum_mode=mode, want=("disp", "left_raw", "right_raw", "lrc"))
    c = e.adcensus_cost(L, R, D, want_ad=True, want_census=True)["cost"]
    e.census9x7(L)
    e.aggregate_cross(p["bgr_left"], c - 0.5, iters=2, mode=SUM_FAST)
    e.scanline(c, L, want=("left", "right", "up", "down", "sum", "disp"))
    a = e.arms_gray(p["gray_left"], 30, variant=0) if H <= W else None
    b = e.arms_gray(p["bgr_left"], 25, variant=1)
    ad = e.ad_cost(p["gray_left"], p["gray_right"], D, 1)
    for mode in (SUM_REFERENCE_ORDER, SUM_FAST):
        e.aggregate_rect(ad, b["left"], b["right"], b["up"], b["down"], variant=2, mode=mode)
    lp, rp = replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)
    e.sad(lp, rp, D, 3, want_costs=True); e.sad(lp, rp, D, 3, right_view=True)
    e.ncc(p["gray_left"], p["gray_right"], D, 4)
    dl = e.asw(lp, rp, D, 3); dr = e.asw(lp, rp, D, 3, right_view=True)
    e.asw_crosscheck(dl, dr)
    d1, d2 = e.wta(c), e.wta(c[:, ::-1].copy())
    e.lrcheck(d1, d2, 2); e.lrcheck_and(d1, d2, 2.0); e.remove_speckles(d1, 1, 30); e.median(d1, 3)
e.close()
print("sanitize_small ok")

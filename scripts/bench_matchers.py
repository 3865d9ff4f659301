"""Throughput of the window matchers and the CrossArm / CBLSM paths at the shapes BASELINE.json names
(configs[0], configs[2]) with the reference's CPU time on a crop beside them.  Development / reporting aid; the
driver-facing benchmark is bench.py.  Prints one JSON object per line."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from oracle import oracle as orc
from stereo_match_traditional_b200.api import SUM_FAST, Engine
from stereo_match_traditional_b200.synth import make_pair, replicate_pad


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    e = Engine(0)
    o = orc.Oracle("ref" if orc.have("ref") else "port")
    dev = "cuda:0"
    out = []
    # configs[0]: SAD 9x9 + WTA, 450x375, D=64
    H, W, D = 375, 450, 64
    p = make_pair(H, W, D, 1001, "scene")
    lp, rp = replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)
    tl, tr = torch.as_tensor(lp, device=dev), torch.as_tensor(rp, device=dev)
    t = gpu_time(lambda: e.sad(tl, tr, D, 3))
    t0 = time.perf_counter(); ref = o.sad(lp, rp, D, 3); tc = time.perf_counter() - t0
    same = bool(np.array_equal(e.sad(lp, rp, D, 3), ref))
    out.append({"path": "SAD GetPointDepthLeft 9x9", "shape": [H, W, D], "gpu_ms": t * 1e3, "gpu_Gcell_per_s": H * W * D / t / 1e9,
                "gpu_Gtap_per_s": H * W * D * 81 / t / 1e9, "cpu_s_full_frame_1core": tc, "speedup": tc / t, "bit_exact": same})
    # configs[2]: NCC 21x21 and ASW 25x25 at 1280x720, D=128
    H, W, D = 720, 1280, 128
    p = make_pair(H, W, D, 1003, "scene")
    gl, gr = torch.as_tensor(p["gray_left"], device=dev), torch.as_tensor(p["gray_right"], device=dev)
    t = gpu_time(lambda: e.ncc(gl, gr, D, 10), reps=3)
    rows = 4
    t0 = time.perf_counter(); o.ncc(p["gray_left"][300 - 10:300 + rows + 10], p["gray_right"][300 - 10:300 + rows + 10], D, 10); tc = time.perf_counter() - t0
    cpu_full = tc * (H - 20) / rows
    out.append({"path": "NCC_algorithem 21x21", "shape": [H, W, D], "gpu_ms": t * 1e3, "gpu_Gcell_per_s": (H - 20) * (W - 20) * D / t / 1e9,
                "cpu_s_full_frame_1core_extrapolated": cpu_full, "cpu_sample": f"{rows} output rows", "speedup": cpu_full / t})
    lp, rp = replicate_pad(p["gray_left"], 12), replicate_pad(p["gray_right"], 12)
    tl, tr = torch.as_tensor(lp, device=dev), torch.as_tensor(rp, device=dev)
    t = gpu_time(lambda: e.asw(tl, tr, D, 11, 50, 30, 40), reps=2)
    rows = 2
    t0 = time.perf_counter(); o.asw(lp[300:300 + rows + 24], rp[300:300 + rows + 24], D, 11, 50, 30, 40); tc = time.perf_counter() - t0
    cpu_full = tc * H / rows
    out.append({"path": "AdaptiveSupportWeight 25x25", "shape": [H, W, D], "gpu_ms": t * 1e3, "gpu_Gcell_per_s": H * W * D / t / 1e9,
                "gpu_Gtap_per_s": H * W * D * 625 / t / 1e9, "cpu_s_full_frame_1core_extrapolated": cpu_full, "cpu_sample": f"{rows} output rows",
                "speedup": cpu_full / t})
    # CBLSM path at Teddy shape: AD + arms + costAggregationV5 x2 + WTA (CBLSM.cpp:64-153, one view)
    H, W, D = 375, 450, 64
    p = make_pair(H, W, D, 1002, "scene")
    gl, gr = torch.as_tensor(p["gray_left"], device=dev), torch.as_tensor(p["gray_right"], device=dev)

    def cblsm():
        ad = e.ad_cost(gl, gr, D, 0)
        a = e.arms_gray(gl, 25, 34, 17, variant=1)
        v1 = e.aggregate_rect(ad, a["left"], a["right"], a["up"], a["down"], variant=2, mode=SUM_FAST)
        v2 = e.aggregate_rect(v1, a["left"], a["right"], a["up"], a["down"], variant=2, mode=SUM_FAST)
        return e.wta(v2)
    t = gpu_time(cblsm, reps=3)
    rows = 40
    gL, gR = p["gray_left"][:rows], p["gray_right"][:rows]
    t0 = time.perf_counter()
    adL, _ = o.cblsm_ad(gL, gR, D); arms = o.cblsm_arms(gL, 25, 34, 17); v = o.cblsm_aggregate_v5(o.cblsm_aggregate_v5(adL, arms), arms); o.cblsm_wta(v)
    tc = (time.perf_counter() - t0) * H / rows
    out.append({"path": "CBLSM AD + ArmLength + costAggregationV5 x2 + WTA (one view, SAT mode)", "shape": [H, W, D], "gpu_ms": t * 1e3,
                "cpu_s_full_frame_1core_extrapolated": tc, "cpu_sample": f"{rows}-row band", "speedup": tc / t})
    for r in out:
        print(json.dumps(r))
    e.close()


if __name__ == "__main__":
    main()

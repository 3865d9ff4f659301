"""How small do aggregated costs get on the benchmark scene?  (decides the domain guard of the fixed-point prefixes)"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine, SUM_FAST
from stereo_match_traditional_b200.synth import make_pair

H, W, D = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1080x1920x256").split("x"))
p = make_pair(H, W, D, 1004, "scene")
dev = "cuda:0"
with Engine(0) as e:
    gl = torch.as_tensor(p["gray_left"].astype(np.float32), device=dev)
    gr = torch.as_tensor(p["gray_right"].astype(np.float32), device=dev)
    bgr = torch.as_tensor(p["bgr_left"], device=dev)
    cost = e.adcensus_cost(gl, gr, D)["cost"]
    for it in (0, 1, 2, 3, 4):
        v = cost if it == 0 else e.aggregate_cross(bgr, cost, iters=it, mode=SUM_FAST)["cost"]
        flag = e.aggregate_fast_status() if it else None
        nz = v[v > 0]
        qs = [float((nz < 2.0 ** -k).sum()) for k in (6, 7, 8, 9, 10, 11, 12, 14, 16)]
        print(f"iters={it} fallback={flag} zeros={int((v == 0).sum())} min_nonzero={float(nz.min()):.3e} "
              f"counts below 2^-k (k=6,7,8,9,10,11,12,14,16): {qs}", flush=True)
        del v, nz

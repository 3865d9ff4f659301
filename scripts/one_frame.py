"""One P_full frame (after one warm-up frame) -- the command profiled with ncu for the launch lists under profiles/."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from stereo_match_traditional_b200.api import Engine, SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.synth import make_pair

H, W, D = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1080x1920x256").split("x"))
mode = SUM_REFERENCE_ORDER if (len(sys.argv) > 2 and sys.argv[2] == "exact") else SUM_FAST
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda:0"
with Engine(0) as e:
    for f in range(frames):
        p = make_pair(H, W, D, 1004 + f, "scene", frame=f)
        t = {k: torch.as_tensor(p[k], device=dev) for k in ("bgr_left", "bgr_right", "gray_left", "gray_right")}
        torch.cuda.synchronize()
        r = e.pipeline(t["bgr_left"], t["bgr_right"], D, t["gray_left"], t["gray_right"], sum_mode=mode)
        e.sync()
        print("frame", f, "redone tasks (left view)", e.aggregate_fast_status(), "valid px", int(torch.isfinite(r["disp"]).sum()), flush=True)
